#!/bin/bash
# Copies the artefacts of the last tools/gpu_round.sh session from gpurun_out/ (scratch) into
# profiles/ (tracked), named per round:  bash tools/collect_profiles.sh "note for the capture"
set -e
cd "$(dirname "$0")/.."
R=r02
cp gpurun_out/r02_parity_errors.json profiles/${R}_parity_errors.json
cp gpurun_out/bench.json profiles/${R}_bench_1gpu.json
[ -f gpurun_out/bench_reference.json ] && cp gpurun_out/bench_reference.json profiles/${R}_bench_reference_arm.json
cp gpurun_out/gpu_tests.log profiles/${R}_gpu_tests.log
cp gpurun_out/reference_tests.log profiles/${R}_reference_suite.log
cp gpurun_out/sanitize_targets.log profiles/${R}_sanitize_targets.log
cp gpurun_out/smoke.log profiles/${R}_smoke.log
if [ -f gpurun_out/r02_prof.ncu-rep ]; then
  ncu -i gpurun_out/r02_prof.ncu-rep --page raw --csv > /tmp/${R}_raw.csv 2>/dev/null
  python tools/ncu_summary.py profiles/${R}_ncu_summary.csv /tmp/${R}_raw.csv
  python tools/ncu_traffic.py /tmp/${R}_raw.csv "${1:-r02 final capture} (gpurun_out/r02_prof.ncu-rep, ncu --set full --clock-control none of bench.py --steps 1 --warmup 1 --no-cpu --no-extras)"
fi
[ -f gpurun_out/r02_launch_list.csv ] && cp gpurun_out/r02_launch_list.csv profiles/${R}_launch_list.csv
ls -la profiles | grep ${R}_
