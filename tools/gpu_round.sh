#!/bin/bash
# One GPU-box session: smoke, GPU tests, parity table, reference suite, bench, sanitizer.
# Usage (from the repo root, under gpurun): bash tools/gpu_round.sh [steps...]
# (`prof` and `launches` each run ncu once, after the same command has run plain: one per session)
mkdir -p gpurun_out
STEPS=${@:-smoke tests parity reftests bench sanitize}
for s in $STEPS; do
  case $s in
    smoke)    timeout 600 python -c "import __graft_entry__ as g; g.build(); g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" ;;
    tests)    timeout 1500 python -m pytest tests -m gpu -q --maxfail=40 --timeout 600 -p no:cacheprovider > gpurun_out/gpu_tests.log 2>&1; echo "tests rc=$?"; tail -n 3 gpurun_out/gpu_tests.log ;;
    parity)   timeout 900 python tools/parity_table.py gpurun_out/r02_parity_errors.json > gpurun_out/parity.log 2>&1; echo "parity rc=$?"; tail -n 5 gpurun_out/parity.log ;;
    reftests) timeout 900 python tools/run_reference_tests.py -q -rfE > gpurun_out/reference_tests.log 2>&1; echo "reftests rc=$?"; tail -n 3 gpurun_out/reference_tests.log ;;
    bench)    timeout 1200 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; head -c 600 gpurun_out/bench.json ;;
    benchref) timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "benchref rc=$?" ;;
    sanitize) bash tools/sanitize.sh gpurun_out/sanitizer; echo "sanitize done" ;;
    prof)     CMD="python bench.py --steps 1 --warmup 1 --no-cpu --no-extras"
              $CMD > gpurun_out/prof_plain.log 2>&1 && \
              ncu --set full --clock-control none --import-source on \
                  -k regex:"fast2|joint_forward_t|joint_dgrad2|joint_wgrad_tc|linear_.*_tc" -c 12 \
                  -f -o gpurun_out/r02_prof $CMD > gpurun_out/prof_ncu.log 2>&1; echo "prof rc=$?"; tail -n 3 gpurun_out/prof_ncu.log ;;
    launches) CMD="python bench.py --steps 2 --warmup 1 --no-cpu --no-extras"
              $CMD > gpurun_out/launches_plain.log 2>&1 && \
              ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
                  --log-file gpurun_out/r02_launch_list.csv $CMD > gpurun_out/launches_ncu.log 2>&1; echo "launches rc=$?" ;;
    benchlib) # A/B of two builds of the library: default vs $LT_AB_LIB (path relative to the repo root)
              timeout 600 python bench.py --steps 10 --warmup 3 --no-extras --no-cpu > gpurun_out/bench_A.json 2> gpurun_out/bench_A.err; echo "bench A rc=$?"
              LT_LIBRARY=$PWD/$LT_AB_LIB timeout 600 python bench.py --steps 10 --warmup 3 --no-extras --no-cpu > gpurun_out/bench_B.json 2> gpurun_out/bench_B.err; echo "bench B rc=$?" ;;
    sanity)   timeout 900 python tools/sanitize_targets.py > gpurun_out/sanitize_targets.log 2>&1; echo "sanity rc=$?"; tail -n 4 gpurun_out/sanitize_targets.log ;;
    benchab)  timeout 600 python bench.py --steps 10 --warmup 3 --no-e2e --no-extras --no-cpu > gpurun_out/bench_norm.json 2> gpurun_out/bench_norm.err; echo "bench norm rc=$?"
              LT_NO_NORM=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-e2e --no-extras --no-cpu > gpurun_out/bench_nonorm.json 2> gpurun_out/bench_nonorm.err; echo "bench nonorm rc=$?" ;;
  esac
done
