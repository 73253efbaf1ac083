#!/usr/bin/env python
"""Benchmark of the lattice hot path: GNAT loss + gradient.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

Workload (BASELINE.json configs[1] with the numerator of configs[3]):
FullNGram(vocab 256, context_size 1) = 257 context states, FrameDependent,
Log-semiring loss = logZ - numerator and its gradient w.r.t. the arc weights,
B = 32 utterances x T = 1000 frames PER GPU (weak scaling), U = 120 labels.

One JSON line is printed by rank 0:
  value  frames*states/s (= N*B*T*C / step time) with the dense arc weights
         already resident in HBM (materialised-weights mode): K1 forward +
         K3 numerator + K2 backward + numerator scatter, through the C ABI.
  e2e    the same metric through the public API the reference exposes,
         `loss = lattice(frames, num_frames, labels, num_labels)` followed by
         `loss.sum().backward()`: frames/labels start in PINNED HOST memory
         (H2D inside the timed region), arc weights come from JointWeightFn
         (H = E = D = 512), parameter gradients are produced (and all-reduced
         over NCCL for N > 1), and the loss is read back to the host.
  roofline      achieved HBM GB/s of the dominant kernel vs the measured peak.
  cpu_baseline  the C/OpenMP port of the reference algorithm (oracle/) timed
                on this box's host cores on a bounded sample (rank 0, N = 1).

  extra_configs  the other BASELINE.json configs, each with ms / step and its
                roofline fraction (configs[2]: trigram FrameLabelDependent(2)
                MaxTropical + Viterbi; bigram FrameLabelDependent(2); a ragged
                batch; the configs[4] per-GPU batch sweep), so that they are
                driver-run numbers.  `--no-extras` skips them.

`--impl reference` times that CPU port through the same frames -> loss -> grads
path (numpy/BLAS joint network + C lattice recursion) and prints the same line
(`config` = the workload the GPU arm runs, `sample_config` = the bounded sample
of it one CPU step covers).
"""

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
  sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = 'lattice loss+grad frames*states/s'
UNIT = 'frames*states/s'


def parse_args():
  ap = argparse.ArgumentParser()
  ap.add_argument('--gpus', type=int, default=1)
  ap.add_argument('--steps', type=int, default=10)
  ap.add_argument('--warmup', type=int, default=3)
  ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
  ap.add_argument('--batch', type=int, default=32, help='utterances per GPU')
  ap.add_argument('--frames', type=int, default=1000)
  ap.add_argument('--vocab', type=int, default=256)
  ap.add_argument('--context-size', type=int, default=1)
  ap.add_argument('--labels', type=int, default=120)
  ap.add_argument('--hidden', type=int, default=512)
  ap.add_argument('--ragged', action='store_true', help='num_frames ~ U{T/2..T}')
  ap.add_argument('--flags', type=int, default=0, help='kernel dispatch flags (last_lattice.h)')
  ap.add_argument('--no-e2e', action='store_true')
  ap.add_argument('--no-extras', action='store_true', help='skip extra_configs')
  ap.add_argument('--no-cpu', action='store_true')
  ap.add_argument('--cpu-seconds', type=float, default=12.0)
  return ap.parse_args()


def peaks():
  path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
  if os.path.exists(path):
    with open(path) as f:
      p = json.load(f)
    return float(p['hbm_gbs']), 'measured (MEASURED_PEAKS.json)'
  return 6650.0, 'fallback (B200_PROFILING.md)'


def config_dict(args, n_gpus):
  c = sum(args.vocab**i for i in range(args.context_size + 1))
  return {
      'workload': (f'GNAT FullNGram(vocab={args.vocab}, context_size={args.context_size}) '
                   f'{c} states, FrameDependent, Log loss+grad, B={args.batch}/GPU T={args.frames} '
                   f'U={args.labels} (BASELINE configs[1] + numerator of configs[3])'),
      'per_gpu_batch': args.batch, 'global_batch': args.batch * n_gpus, 'frames': args.frames,
      'states': c, 'vocab': args.vocab, 'labels': args.labels,
      'joint_hidden': args.hidden, 'ragged': bool(args.ragged),
      'parallelism': f'utterance-sharded x{n_gpus}',
      'l2': 'inputs (8.4 GB of arc weights per pass) are larger than the 126 MB L2',
  }


# ---------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------

class ClockSampler:
  """Polls NVML (SM clock, max SM clock, power, throttle reasons) every 20 ms
  on a background thread while the timed region runs."""
  REASONS = {
      'hw_slowdown': 0x8, 'sw_thermal_slowdown': 0x20, 'hw_thermal_slowdown': 0x40,
      'sw_power_cap': 0x4, 'hw_power_brake_slowdown': 0x80,
  }

  def __init__(self, index):
    self.index = index
    self.samples = []
    self.stop_flag = False
    self.thread = None
    self.error = None

  def _run(self):
    try:
      import pynvml
      pynvml.nvmlInit()
      visible = os.environ.get('CUDA_VISIBLE_DEVICES')
      index = self.index
      if visible:
        try:
          index = int(visible.split(',')[self.index])
        except ValueError:
          pass
      h = pynvml.nvmlDeviceGetHandleByIndex(index)
      smax = pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)
      while not self.stop_flag:
        sm = pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)
        try:
          reasons = pynvml.nvmlDeviceGetCurrentClocksEventReasons(h)
        except Exception:
          reasons = pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)
        try:
          power = pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0
        except Exception:
          power = None
        self.samples.append((time.time(), sm, smax, reasons, power))
        time.sleep(0.02)
    except Exception as e:  # pragma: no cover
      self.error = repr(e)

  def start(self):
    self.thread = threading.Thread(target=self._run, daemon=True)
    self.thread.start()

  def stop(self, t0, t1):
    self.stop_flag = True
    if self.thread is not None:
      self.thread.join(timeout=2.0)
    inside = [x for x in self.samples if t0 <= x[0] <= t1] or self.samples
    if not inside:
      return {'sm_mhz': None, 'sm_max_mhz': None, 'samples': 0,
              'reasons': [f'nvml unavailable: {self.error}']}
    mask = 0
    for x in inside:
      mask |= int(x[3])
    reasons = sorted(k for k, bit in self.REASONS.items() if mask & bit)
    powers = [x[4] for x in inside if x[4] is not None]
    return {'sm_mhz': statistics.median(x[1] for x in inside), 'sm_max_mhz': inside[0][2],
            'samples': len(inside), 'reasons': reasons,
            'power_w_max': max(powers) if powers else None}


# ---------------------------------------------------------------------------
# CPU port (oracle) legs
# ---------------------------------------------------------------------------

def cpu_sample_inputs(args, b, t, seed=0):
  rng = np.random.RandomState(seed)
  v, h = args.vocab, args.hidden
  c = sum(v**i for i in range(args.context_size + 1))
  u = min(args.labels, max(1, t - 1))
  return dict(
      frames=rng.randn(b, t, h).astype(np.float32), cache=rng.randn(c, h).astype(np.float32),
      w_ctx=(rng.randn(h, h) * 0.3 / np.sqrt(h)).astype(np.float32),
      w_frame=(rng.randn(h, h) * 0.3 / np.sqrt(h)).astype(np.float32),
      w_blank=(rng.randn(h) * 0.3).astype(np.float32), b_blank=np.float32(0.1),
      w_vocab=(rng.randn(v, h) * 0.3).astype(np.float32),
      b_vocab=(rng.randn(v) * 0.1).astype(np.float32),
      num_frames=np.full([b], t, np.int32), labels=rng.randint(1, v + 1, (b, u)).astype(np.int32),
      num_labels=np.full([b], u, np.int32), c=c, u=u)


def cpu_step(args, s, with_joint):
  """One loss+grad pass of the CPU port.  with_joint=True is the frames -> loss
  -> parameter-gradients path (numpy/BLAS JointWeightFn + C lattice recursion);
  False starts from materialised arc weights (lattice only)."""
  from oracle import c_oracle
  v = args.vocab
  if with_joint:
    pc = s['cache'] @ s['w_ctx'].T
    pf = s['frames'] @ s['w_frame'].T
    joint = np.tanh(pc[None, None] + pf[:, :, None, :])                 # [B,T,C,H]
    lexical = joint @ s['w_vocab'].T + s['b_vocab']
    blank = joint @ s['w_blank'] + s['b_blank']
  else:
    blank, lexical = s['blank'], s['lexical']
  loss, gb, gl, _, _ = c_oracle.lattice_loss_and_grads(
      blank, lexical, s['num_frames'], s['labels'], s['num_labels'], v, args.context_size, -1)
  if with_joint:
    g_joint = gl @ s['w_vocab'] + gb[..., None] * s['w_blank']
    g_pre = g_joint * (1 - joint * joint)
    h = joint.shape[-1]
    g_wv = gl.reshape(-1, v).T @ joint.reshape(-1, h)
    g_wb = gb.reshape(-1) @ joint.reshape(-1, h)
    g_pc = g_pre.sum((0, 1))
    g_pf = g_pre.sum(2)
    g_wctx = g_pc.T @ s['cache']
    g_wframe = g_pf.reshape(-1, h).T @ s['frames'].reshape(-1, h)
    return float(loss.sum()) + 0 * float(g_wv[0, 0] + g_wb[0] + g_wctx[0, 0] + g_wframe[0, 0])
  return float(loss.sum())


def cpu_bounded_run(args, with_joint, seconds, reps=1):
  """Sizes a sample (same widths, fewer utterances/frames) so that one pass takes
  about `seconds`, and times `reps` passes.  Returns (units/s, description)."""
  from oracle import c_oracle
  import __graft_entry__ as ge
  ge.build_oracle()
  # use every host core, also under torchrun (which exports OMP_NUM_THREADS=1)
  cores = len(os.sched_getaffinity(0)) if hasattr(os, 'sched_getaffinity') else os.cpu_count()
  c_oracle.set_threads(cores)
  try:
    import threadpoolctl
    threadpoolctl.threadpool_limits(limits=cores)
  except Exception:
    pass
  threads = c_oracle.num_threads()
  b = max(1, min(args.batch, threads))
  t = 8
  s = cpu_sample_inputs(args, b, t)
  if not with_joint:
    rng = np.random.RandomState(1)
    s['blank'] = rng.randn(b, t, s['c']).astype(np.float32)
    s['lexical'] = rng.randn(b, t, s['c'], args.vocab).astype(np.float32)
  cpu_step(args, s, with_joint)       # warm-up (page faults, thread pool)
  t0 = time.perf_counter()
  cpu_step(args, s, with_joint)
  probe = time.perf_counter() - t0
  t_full = int(max(8, min(args.frames, t * seconds / max(probe, 1e-4))))
  s = cpu_sample_inputs(args, b, t_full)
  if not with_joint:
    rng = np.random.RandomState(1)
    s['blank'] = rng.randn(b, t_full, s['c']).astype(np.float32)
    s['lexical'] = rng.randn(b, t_full, s['c'], args.vocab).astype(np.float32)
  times = []
  for _ in range(reps):
    t0 = time.perf_counter()
    cpu_step(args, s, with_joint)
    times.append(time.perf_counter() - t0)
  dt = statistics.median(times)
  units = b * t_full * s['c']
  desc = (f'{b} utterances x {t_full} frames x {s["c"]} states, U={s["u"]}, '
          f'{"frames->JointWeightFn(numpy/BLAS)->" if with_joint else ""}lattice loss+grad '
          f'(C/OpenMP port of the reference algorithm), {dt:.2f} s per pass')
  sample = {'utterances': b, 'frames': t_full, 'states': s['c'], 'labels': s['u'],
            'frames_x_states_per_step': units}
  return units / dt, threads, desc, dt, sample


def run_reference(args):
  rank = int(os.environ.get('RANK', '0'))
  if rank != 0:
    return
  steps = max(1, args.steps)
  per_step = min(args.cpu_seconds, 120.0 / (steps + max(args.warmup, 0) + 1))
  val, threads, desc, dt, sample = cpu_bounded_run(args, with_joint=True, seconds=per_step,
                                                   reps=steps)
  line = {
      'impl': 'reference', 'metric': METRIC, 'value': val, 'unit': UNIT, 'n_gpus': args.gpus,
      'steps': steps, 'warmup': args.warmup, 'ms_per_step': dt * 1e3, 'higher_is_better': True,
      'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
      'config': config_dict(args, args.gpus),
      # one CPU "step" is a bounded SAMPLE of that workload (same widths, fewer utterances and
      # frames); ms_per_step is the time of one pass over the sample, value its rate
      'sample_config': sample,
      'cpu_baseline': {'value': val, 'unit': UNIT, 'cores': threads, 'kind': 'port',
                       'sample': desc},
      'e2e': {'value': val, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
      'note': ('the reference is pure Python/PyTorch and is not present on the GPU box; this is '
               'the oracle port of its algorithm (oracle/lattice_oracle.c + numpy joint network) '
               'on the host cores'),
  }
  print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------

def kernel_name(name):
  """C-ABI entry point -> the name the kernels are reported under (the *_norm / *_checked
  variants are the same kernels with extra outputs)."""
  return name.replace('_norm', '').replace('_checked', '')


def ncu_traffic(kernel):
  """dram__bytes_read.sum + dram__bytes_write.sum per launch of `kernel` from the committed
  `ncu --set full` capture of this workload (profiles/r02_ncu_traffic.json), or (None, why) when
  the kernel's source has changed since the capture."""
  import hashlib
  path = os.path.join(ROOT, 'profiles', 'r02_ncu_traffic.json')
  if not os.path.exists(path):
    return None, 'no committed capture (profiles/r02_ncu_traffic.json)'
  with open(path) as f:
    doc = json.load(f)
  ent = doc.get('kernels', {}).get(kernel)
  if not ent:
    return None, f'{kernel} not in profiles/r02_ncu_traffic.json'
  sources = ent['source'] if isinstance(ent['source'], list) else [ent['source']]
  h = hashlib.sha256()
  for src in sources:
    with open(os.path.join(ROOT, src), 'rb') as f:
      h.update(f.read())
  if h.hexdigest() != ent['source_sha256']:
    return None, (f'{", ".join(sources)} changed since the capture '
                  f'({doc.get("capture", "profiles/")}): re-profile')
  return float(ent['dram_bytes_per_launch']), doc.get('capture', 'profiles/r02_ncu_traffic.json')


def run_b200(args):
  import torch
  import torch.distributed as dist
  import __graft_entry__ as ge

  world = int(os.environ.get('WORLD_SIZE', '1'))
  rank = int(os.environ.get('RANK', '0'))
  local_rank = int(os.environ.get('LOCAL_RANK', '0'))
  torch.cuda.set_device(local_rank)
  dev = torch.device('cuda', local_rank)
  if world > 1:
    os.environ.setdefault('NCCL_DEBUG', 'WARN')    # keep NCCL's banner off stdout
    dist.init_process_group('nccl', device_id=dev)
  if rank == 0:
    ge.build()
  if world > 1:
    dist.barrier()

  import last_torch_b200 as last_torch
  from last_torch_b200 import _native as N
  from last_torch_b200 import distributed as ltdist
  from last_torch_b200 import ops

  lib = N.lib()
  B, T, V, n, U, H = args.batch, args.frames, args.vocab, args.context_size, args.labels, args.hidden
  context = last_torch.contexts.FullNGram(vocab_size=V, context_size=n)
  C = context.num_states()
  gen = torch.Generator(device=dev).manual_seed(1234 + rank)
  cpu_gen = torch.Generator().manual_seed(1234 + rank)
  peak, peak_src = peaks()

  def barrier_sync():
    if world > 1:
      dist.barrier()
    torch.cuda.synchronize()

  def timed(step_fn, steps, warmup, sample_clocks=False):
    for _ in range(warmup):
      step_fn()
    barrier_sync()
    launches0 = lib.lt_launch_count()
    timer = []
    N.KERNEL_TIMER = timer
    sampler = ClockSampler(local_rank) if sample_clocks else None
    if sampler:
      sampler.start()
      time.sleep(0.25)
    start = torch.cuda.Event(enable_timing=True)
    end = torch.cuda.Event(enable_timing=True)
    w0 = time.time()
    start.record()
    for _ in range(steps):
      step_fn()
    end.record()
    barrier_sync()
    w1 = time.time()
    N.KERNEL_TIMER = None
    clocks = sampler.stop(w0, w1) if sampler else None
    ms = torch.tensor([start.elapsed_time(end)], device=dev)
    if world > 1:
      dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    launches = lib.lt_launch_count() - launches0
    kernels = {}
    for name, s, e in timer:
      kernels.setdefault(kernel_name(name), []).append(s.elapsed_time(e))
    return float(ms.item()) / steps, launches, kernels, clocks

  def all_sum(x):
    t = torch.tensor([float(x)], device=dev, dtype=torch.float64)
    if world > 1:
      dist.all_reduce(t)
    return float(t.item())

  def lattice_inputs(b, t, v, nn, u, ragged):
    c = sum(v**i for i in range(nn + 1))
    if ragged:
      num_frames = torch.randint(t // 2, t + 1, [b], generator=cpu_gen).to(torch.int32)
    else:
      num_frames = torch.full([b], t, dtype=torch.int32)
    labels = torch.randint(1, v + 1, [b, u], generator=cpu_gen).to(torch.int32)
    num_labels = torch.full([b], u, dtype=torch.int32)
    blank = torch.randn([b, t, c], device=dev, generator=gen).requires_grad_()
    lexical = torch.randn([b, t, c, v], device=dev, generator=gen).requires_grad_()
    nl_d = num_labels.to(dev)
    states, next_labels, _ = ops.walk_states(labels.to(dev), nl_d, v, nn)
    return dict(c=c, num_frames=num_frames, nf_d=num_frames.to(dev), nl_d=nl_d, labels=labels,
                num_labels=num_labels, blank=blank, lexical=lexical, states=states,
                next_labels=next_labels)

  def loss_grad_step(x, v, nn, k, flags):
    loss, _, _ = ops.LatticeLoss.apply(x['blank'], x['lexical'], x['nf_d'], x['states'],
                                       x['next_labels'], x['nl_d'], v, nn, k, flags)
    total = loss.sum()
    gb, gl = torch.autograd.grad(total, (x['blank'], x['lexical']))
    return total, gb, gl

  def kernel_table(kernels, steps, w_bytes):
    kern = {}
    for name, ts in kernels.items():
      kern[name] = {'ms': statistics.mean(ts), 'calls_per_step': len(ts) / steps}
    alg = {'lt_lattice_forward': 1.0 * w_bytes, 'lt_lattice_backward': 2.0 * w_bytes}
    for name, nbytes in alg.items():
      if name in kern:
        kern[name]['algorithmic_gb'] = nbytes / 1e9
        kern[name]['gbps'] = nbytes / 1e9 / (kern[name]['ms'] * 1e-3)
        kern[name]['frac'] = kern[name]['gbps'] / peak
    return kern, alg

  # ---- device-resident arm: dense arc weights already in HBM -----------------
  x = lattice_inputs(B, T, V, n, U, args.ragged)
  num_frames = x['num_frames']
  loss_sum = torch.zeros([1], device=dev)

  def resident_step():
    total, gb, gl = loss_grad_step(x, V, n, -1, args.flags)
    if world > 1:
      loss_sum.copy_(total.detach().reshape(1))
      dist.all_reduce(loss_sum)       # the only exchange in materialised-weights mode
    return gb, gl

  ms_step, launches, kernels, clocks = timed(resident_step, args.steps, args.warmup,
                                             sample_clocks=True)
  # frames*states actually processed (padding frames of a ragged batch are not counted)
  units = all_sum(float(num_frames.sum())) * C
  value = units / (ms_step * 1e-3)
  w_bytes = int(num_frames.sum()) * C * (V + 1) * 4      # arc weights of this rank's real frames
  kern, alg = kernel_table(kernels, args.steps, w_bytes)
  dom = max((k for k in kern if k in alg), key=lambda k: kern[k]['ms'], default=None)
  roofline = None
  if dom:
    default_shape = (B, T, V, n, U) == (32, 1000, 256, 1, 120) and args.flags == 0 and not args.ragged
    traffic, traffic_src = ncu_traffic(dom) if default_shape else (None, 'non-default shape')
    roofline = {'bound': 'hbm', 'kernel': dom, 'achieved': kern[dom]['gbps'], 'peak': peak,
                'peak_source': peak_src, 'unit': 'GB/s', 'frac': kern[dom]['frac'],
                'traffic': traffic, 'traffic_source': traffic_src,
                'algorithmic_bytes_per_launch': alg[dom],
                'whole_step': {'algorithmic_gb': 3.0 * w_bytes / 1e9,
                               'gbps': 3.0 * w_bytes / 1e9 / (ms_step * 1e-3),
                               'frac': 3.0 * w_bytes / 1e9 / (ms_step * 1e-3) / peak}}
  del x
  torch.cuda.empty_cache()

  # ---- end-to-end arm: public API, host buffers -------------------------------
  e2e = None
  if not args.no_e2e:
    torch.manual_seed(4321)       # same parameters on every rank
    lattice = last_torch.RecognitionLattice(
        context=context, alignment=last_torch.alignments.FrameDependent(),
        weight_fn_cacher_factory=lambda c: last_torch.weight_fns.SharedEmbCacher(
            num_context_states=c.shape()[0], embedding_size=H, device=str(dev)),
        weight_fn_factory=lambda c: last_torch.weight_fns.JointWeightFn(
            vocab_size=c.shape()[1], hidden_size=H, device=str(dev), embedding_size=H,
            feature_size=H))
    lattice.kernel_flags = args.flags
    labels = torch.randint(1, V + 1, [B, U], generator=cpu_gen).to(torch.int32)
    host = dict(frames=torch.randn([B, T, H], generator=cpu_gen).pin_memory(),
                num_frames=num_frames.pin_memory(), labels=labels.pin_memory(),
                num_labels=torch.full([B], U, dtype=torch.int32).pin_memory())
    loss_h = torch.empty([B], dtype=torch.float32).pin_memory()
    h2d = sum(t.numel() * t.element_size() for t in host.values())
    d2h = loss_h.numel() * 4
    # Double-buffered inputs: the host-to-device copy of step i+1 is issued on a copy stream
    # BEFORE step i's kernels and overlaps them; every timed step issues (and pays for) exactly
    # one copy of a full input batch.
    cur = torch.cuda.current_stream(dev)
    copy_stream = torch.cuda.Stream(device=dev)
    slots = [{k: torch.empty_like(v, device=dev) for k, v in host.items()} for _ in range(2)]
    ready = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    begun = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    consumed = [torch.cuda.Event() for _ in range(2)]
    for ev in consumed:
      ev.record(cur)
    state = {'i': 0, 'primed': False, 'h2d_ms': [], 'ar_ms': []}

    def issue_copy(slot):
      with torch.cuda.stream(copy_stream):
        copy_stream.wait_event(consumed[slot])
        begun[slot].record(copy_stream)
        for k, v in host.items():
          slots[slot][k].copy_(v, non_blocking=True)
        ready[slot].record(copy_stream)

    def e2e_step():
      slot = state['i'] & 1
      if not state['primed']:
        issue_copy(slot)
        state['primed'] = True
      issue_copy(slot ^ 1)                      # inputs of the NEXT step, under this step's kernels
      cur.wait_event(ready[slot])
      d = slots[slot]
      if world > 1:
        a0 = torch.cuda.Event(enable_timing=True)
        a1 = torch.cuda.Event(enable_timing=True)
        total, grads, loss = ltdist.local_loss_and_grads(
            lattice, d['frames'], d['num_frames'], d['labels'], d['num_labels'])
        consumed[slot].record(cur)
        a0.record(cur)
        total, grads = ltdist.all_reduce_loss_and_grads(total, grads)   # NCCL over NVLink
        a1.record(cur)
      else:
        loss = lattice(frames=d['frames'], num_frames=d['num_frames'], labels=d['labels'],
                       num_labels=d['num_labels'])
        grads = torch.autograd.grad(loss.sum(), list(lattice.parameters()))
        consumed[slot].record(cur)
      loss_h.copy_(loss.detach(), non_blocking=True)
      cur.synchronize()
      if N.KERNEL_TIMER is not None:
        state['h2d_ms'].append(begun[slot].elapsed_time(ready[slot]))
        if world > 1:
          state['ar_ms'].append(a0.elapsed_time(a1))
      state['i'] += 1
      return loss_h

    e_steps, e_warm = max(1, args.steps), max(1, args.warmup)
    ms_e2e, e_launches, e_kernels, e_clocks = timed(e2e_step, e_steps, e_warm, sample_clocks=True)
    kms = {k: statistics.mean(v) for k, v in e_kernels.items()}
    kms['lt_h2d'] = statistics.mean(state['h2d_ms']) if state['h2d_ms'] else None
    if world > 1:
      kms['nccl_all_reduce'] = statistics.mean(state['ar_ms']) if state['ar_ms'] else None
    e2e = {'value': units / (ms_e2e * 1e-3), 'unit': UNIT, 'ms_per_step': ms_e2e,
           'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h, 'steps': e_steps,
           'warmup': e_warm, 'gpu_launches_per_step': e_launches / e_steps,
           'kernels_ms': kms, 'clocks': e_clocks,
           'h2d': ('double-buffered on a copy stream: the copy of step i+1 is issued before step '
                   'i\'s kernels; lt_h2d is its duration on that stream (overlapped), '
                   'nccl_all_reduce the flat loss + parameter-gradient all-reduce'),
           'api': 'RecognitionLattice.forward + autograd.grad w.r.t. JointWeightFn/SharedEmbCacher '
                  'parameters' + ('; distributed.local_loss_and_grads + all_reduce_loss_and_grads'
                                  if world > 1 else ''),
           'grad_handover': ('split rows (bf16 hi | lo) from the lattice backward to the joint '
                             'backward (ops.JointLatticeLoss)' if lattice.split_grad_handover
                             else 'float32')}
    del lattice, slots
    torch.cuda.empty_cache()

  # ---- the other BASELINE configs, each as a short timed run --------------------------------
  extras = None
  if not args.no_extras:
    extras = []
    x_steps, x_warm = 3, 3

    def add(name, fn, b, t, v, nn, passes, frames_done=None, note=None):
      """passes: how many times W (the fp32 arc weights) crosses HBM per step."""
      c = sum(v**i for i in range(nn + 1))
      try:
        ms, _, kk, _ = timed(fn, x_steps, x_warm)
      except Exception as e:     # an extra must never take the headline line down
        extras.append({'config': name, 'error': repr(e)[:300]})
        return
      frames = all_sum(frames_done if frames_done is not None else b * t)
      wb = frames * c * (v + 1) * 4 / max(world, 1)        # per rank
      ent = {'config': name, 'per_gpu_batch': b, 'frames': t, 'states': c, 'vocab': v, 'ms': ms,
             'frames_states_per_s': frames * c / (ms * 1e-3),
             'roofline': {'bound': 'hbm', 'algorithmic_gb': passes * wb / 1e9,
                          'gbps': passes * wb / 1e9 / (ms * 1e-3),
                          'frac': passes * wb / 1e9 / (ms * 1e-3) / peak, 'peak': peak},
             'kernels_ms': {k: statistics.mean(tt) for k, tt in kk.items()}}
      if note:
        ent['note'] = note
      extras.append(ent)

    def run_extra(name, b, t, v, nn, k, u, mode, ragged=False, note=None):
      try:
        xx = lattice_inputs(b, t, v, nn, u, ragged)
      except Exception as e:
        extras.append({'config': name, 'error': repr(e)[:300]})
        return
      frames_done = float(xx['num_frames'].sum())
      if mode == 'lossgrad':
        add(name, lambda: loss_grad_step(xx, v, nn, k, 0), b, t, v, nn, 3.0, frames_done, note)
      elif mode == 'forward':
        bl, lx = xx['blank'].detach(), xx['lexical'].detach()
        add(name, lambda: ops._lattice_forward_raw(N.LOG, v, nn, k, bl, lx, xx['nf_d'], 0, False,
                                                   False, norm=True),
            b, t, v, nn, 1.0, frames_done, note)
      elif mode == 'entropy':
        bl, lx = xx['blank'].detach(), xx['lexical'].detach()
        add(name, lambda: ops.lattice_expectation(bl, lx, xx['nf_d'], v, nn, k), b, t, v, nn, 2.0,
            frames_done, note)
      else:     # MaxTropical shortest distance + Viterbi back-trace
        bl, lx = xx['blank'].detach(), xx['lexical'].detach()
        add(name, lambda: ops.viterbi_path(bl, lx, xx['nf_d'], v, nn, k), b, t, v, nn, 1.0,
            frames_done, note)
      del xx
      torch.cuda.empty_cache()

    # configs[4]: the per-GPU batch sweep (weak scaling: N x these under torchrun)
    for bb in (64, 128):
      run_extra(f'configs[4] bigram vocab 256 T=1000 Log loss+grad, B={bb}/GPU '
                f'(global {bb * world})', bb, 1000, 256, 1, -1, 120, 'lossgrad')
    if world == 1:
      run_extra('configs[2] trigram vocab 64 (4161 states) FrameLabelDependent(2) MaxTropical '
                'shortest distance + Viterbi, B=32 T=500', 32, 500, 64, 2, 2, 60, 'viterbi')
      run_extra('configs[2] geometry, FrameDependent MaxTropical + Viterbi, B=32 T=500',
                32, 500, 64, 2, -1, 60, 'viterbi')
      run_extra('configs[2] geometry, FrameLabelDependent(2) Log loss+grad, B=32 T=500',
                32, 500, 64, 2, 2, 60, 'lossgrad')
      run_extra('bigram vocab 256 FrameLabelDependent(2) Log loss+grad at configs[1] geometry, '
                'B=32 T=1000', 32, 1000, 256, 1, 2, 120, 'lossgrad')
      run_extra('bigram vocab 256 FrameLabelDependent(2) MaxTropical shortest distance + '
                'Viterbi at configs[1] geometry, B=32 T=1000', 32, 1000, 256, 1, 2, 120, 'viterbi')
      run_extra('configs[1] ragged: num_frames ~ U{T/2..T}, Log loss+grad, B=32 T=1000 '
                '(real frames counted)', 32, 1000, 256, 1, -1, 120, 'lossgrad', ragged=True)
      run_extra('configs[4] ragged: num_frames ~ U{T/2..T}, Log loss+grad, B=96/GPU (utterances '
                'are handed to clusters longest first)', 96, 1000, 256, 1, -1, 120, 'lossgrad',
                ragged=True)
      run_extra('configs[1] Log loss+grad, B=48/GPU', 48, 1000, 256, 1, -1, 120, 'lossgrad')
      run_extra('configs[1] Log forward only, B=32/GPU', 32, 1000, 256, 1, -1, 120, 'forward')
      run_extra('configs[1] Log forward only, B=8/GPU', 8, 1000, 256, 1, -1, 120, 'forward')
      run_extra('configs[1] geometry, path entropy (expectation semiring by forward-backward, no '
                'posterior tensor: 2 passes over W), B=32 T=1000', 32, 1000, 256, 1, -1, 120,
                'entropy')

    if world == 1:
      # north_star (4): JointWeightFn fused into the recursion vs joint kernel -> HBM -> K1, at the
      # shape where recomputing the logits is cheapest (vocab 64, hidden 128): Viterbi decoding
      try:
        fv, fh = 64, 128
        torch.manual_seed(99)
        flat = last_torch.RecognitionLattice(
            context=last_torch.contexts.FullNGram(vocab_size=fv, context_size=1),
            alignment=last_torch.alignments.FrameDependent(),
            weight_fn_cacher_factory=lambda c: last_torch.weight_fns.SharedEmbCacher(
                num_context_states=c.shape()[0], embedding_size=fh, device=str(dev)),
            weight_fn_factory=lambda c: last_torch.weight_fns.JointWeightFn(
                vocab_size=c.shape()[1], hidden_size=fh, device=str(dev), embedding_size=fh,
                feature_size=fh))
        fx = torch.randn([32, 1000, fh], device=dev, generator=gen)
        fnf = torch.full([32], 1000, dtype=torch.int32, device=dev)
        ent = {'config': ('north_star (4): JointWeightFn fused into the forward recursion vs joint '
                          'kernel -> HBM -> K1; bigram vocab 64, hidden 128, B=32 T=1000, '
                          'MaxTropical shortest_path (Viterbi decoding)'),
               'logits_gb_unfused': 32 * 1000 * 65 * 65 * 4 / 1e9}
        for fused in (False, True):
          flat.fused_inference = fused
          ms, _, kk, _ = timed(lambda: flat.shortest_path(frames=fx, num_frames=fnf), x_steps, x_warm)
          key = 'fused' if fused else 'unfused'
          ent[key + '_ms'] = ms
          ent[key + '_kernels_ms'] = {k: statistics.mean(tt) for k, tt in kk.items()}
        extras.append(ent)
        del flat, fx
        torch.cuda.empty_cache()
      except Exception as e:
        extras.append({'config': 'north_star (4) fused inference', 'error': repr(e)[:300]})

  cpu = None
  if rank == 0 and world == 1 and not args.no_cpu:
    val, threads, desc, _, _ = cpu_bounded_run(args, with_joint=False, seconds=args.cpu_seconds)
    cpu = {'value': val, 'unit': UNIT, 'cores': threads, 'kind': 'port', 'sample': desc}

  if rank == 0:
    line = {
        'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': ms_step, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': config_dict(args, world), 'clocks': clocks, 'gpu_launches': int(launches),
        'roofline': roofline, 'kernels': kern, 'e2e': e2e, 'cpu_baseline': cpu,
        'extra_configs': extras,
    }
    print(json.dumps(line), flush=True)
  if world > 1:
    dist.barrier()
    dist.destroy_process_group()


def main():
  args = parse_args()
  # stdout carries exactly ONE line, the JSON result: everything else that native libraries
  # write to file descriptor 1 (NCCL's version banner, for one) is sent to stderr instead.
  sys.stdout.flush()
  real_stdout = os.fdopen(os.dup(1), 'w')
  os.dup2(2, 1)
  sys.stdout = real_stdout
  if args.impl == 'reference':
    run_reference(args)
  else:
    run_b200(args)
  real_stdout.flush()


if __name__ == '__main__':
  main()
