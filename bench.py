#!/usr/bin/env python
"""bench.py -- CenterTrack per-frame inference hot path on B200 (contract: see DESIGN.md section 6).

  python bench.py --gpus N --steps K --warmup W [--batch B] [--impl reference]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[1]): DLA-34 coco_tracking, 512x512, bf16, synthetic frame pairs +
pre_hm, K=100.  One step = the hot path (network + fused sigmoid + fused decode) over one batch of B
frames per GPU.  metric = frames/sec, whole job.  Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np
import torch

METRIC = 'frames/sec (device-timed) DLA-34 512x512'
GFLOP_PER_FRAME = 72.56          # algorithmic, BASELINE.md section 2 (coco_tracking 512x512)
H = W = 512
K = 100


def _peaks():
  p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
  if os.path.exists(p):
    d = json.load(open(p))
    return d.get('bf16_tflops_sustained', d.get('bf16_tflops')), d.get('hbm_gbs'), 'measured'
  return 1400.0, 6650.0, 'fallback'


class ClockSampler(threading.Thread):
  """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
  Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,'
       'clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
       'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

  def __init__(self, gpu):
    super().__init__(daemon=True)
    self.gpu, self.rows, self.stop_flag = gpu, [], False

  def run(self):
    while not self.stop_flag:
      try:
        r = subprocess.run(['nvidia-smi', '-i', str(self.gpu), '--query-gpu=' + self.Q,
                            '--format=csv,noheader,nounits'], capture_output=True, text=True, timeout=5)
        if r.returncode == 0 and r.stdout.strip():
          self.rows.append([c.strip() for c in r.stdout.strip().split(',')])
      except Exception:
        pass
      time.sleep(0.1)

  def summary(self):
    self.stop_flag = True
    sm = [float(r[1]) for r in self.rows if r[1].replace('.', '').isdigit()]
    mx = [float(r[2]) for r in self.rows if r[2].replace('.', '').isdigit()]
    reasons = set()
    for r in self.rows:
      for name, v in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), r[5:9]):
        if v.lower().startswith('active'):
          reasons.add(name)
    return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': max(mx) if mx else None,
            'reasons': sorted(reasons), 'samples': len(self.rows)}


def _use_host_threads():
  """torch's own default thread count is kept (64 on the GPU box: oversubscribing its logical CPUs made the oracle
  several times slower); only when a launcher pinned it to one thread (torchrun exports OMP_NUM_THREADS=1) is it
  raised, to half the logical CPUs (= physical cores on an SMT-2 host).  CTB_CPU_THREADS overrides."""
  n = int(os.environ.get('CTB_CPU_THREADS', '0'))
  if n <= 0 and torch.get_num_threads() <= 1:
    n = max(1, (os.cpu_count() or 2) // 2)
  if n > 0:
    torch.set_num_threads(n)


def _oracle_step(n_frames=1, budget_s=None):
  """The CPU restatement of the reference path (oracle/), timed on this host: network + sigmoid +
  decode for up to n_frames 512x512 frame pairs (stops early once budget_s seconds have elapsed, so a slow or
  oversubscribed host cannot stall the bench).  Returns (seconds, threads, frames done)."""
  sys.path.insert(0, os.path.join(ROOT, 'oracle'))
  sys.path.insert(0, os.path.join(ROOT, 'tests'))
  import ct_oracle as co
  from centertrack_b200 import synthetic as wt
  from helpers import make_model
  opt, model, sd = make_model('coco_tracking')
  orc = co.DLA34Oracle(sd, opt.heads)
  img, pre, hm = wt.synthetic_inputs(1, H, W)
  done = 0
  t0 = time.perf_counter()
  for _ in range(n_frames):
    out = co.sigmoid_output(orc.forward(img, pre, hm))
    co.generic_decode(out, K)
    done += 1
    if budget_s is not None and time.perf_counter() - t0 > budget_s:
      break
  return time.perf_counter() - t0, torch.get_num_threads(), done


def run_reference(args, rank, world):
  """--impl reference: the reference's CPU implementation of the path (oracle port; the Python
  reference itself cannot travel to the GPU box) on all host threads.  Rank 0 only."""
  if rank != 0:
    return
  _use_host_threads()
  frames_per_step = 1
  for _ in range(min(args.warmup, 1)):
    _oracle_step(1)
  t, thr = 0.0, 1
  steps = min(args.steps, 6)
  done = 0
  for _ in range(steps):
    dt, thr, _n = _oracle_step(frames_per_step)
    t += dt
    done += 1
    if t > 60.0:                                        # bounded sample on any host
      break
  steps = done
  fps = steps * frames_per_step / t
  line = {'impl': 'reference', 'metric': METRIC, 'value': fps, 'unit': 'frames/s', 'n_gpus': args.gpus,
          'steps': steps, 'warmup': min(args.warmup, 1), 'ms_per_step': 1000 * t / steps,
          'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32',
          'data': 'synthetic',
          'config': {'workload': 'DLA-34 coco_tracking 512x512 frame pairs + pre_hm, K=100 (BASELINE configs[1])',
                     'frames_per_step': frames_per_step},
          'cpu_baseline': {'value': fps, 'unit': 'frames/s', 'cores': thr, 'kind': 'port',
                           'sample': '%d steps x %d frame (oracle/ct_oracle.py: torch-CPU fp32 convs + '
                                     'restated DCNv2/decode)' % (steps, frames_per_step)},
          'e2e': {'value': fps, 'unit': 'frames/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}
  _emit(line)


_REAL_STDOUT = None


def _emit(line):
  data = (json.dumps(line) + '\n').encode()
  if _REAL_STDOUT is None:
    sys.stdout.write(data.decode())
    sys.stdout.flush()
  else:
    sys.stdout.flush()
    os.write(_REAL_STDOUT, data)


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument('--gpus', type=int, default=1)
  ap.add_argument('--steps', type=int, default=30)
  ap.add_argument('--warmup', type=int, default=5)
  ap.add_argument('--batch', type=int, default=32, help='frames (independent streams) per GPU per step')
  ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
  ap.add_argument('--precision', default='bf16', choices=['bf16', 'fp32'])
  ap.add_argument('--no-cpu-baseline', action='store_true')
  args = ap.parse_args()
  # stdout carries exactly ONE JSON line: everything any library writes to fd 1 during the run (NCCL prints a version
  # banner there) is sent to stderr, and the result line goes to the saved descriptor (see _emit)
  global _REAL_STDOUT
  sys.stdout.flush()
  _REAL_STDOUT = os.dup(1)
  os.dup2(2, 1)
  args.warmup = max(args.warmup, 3)

  rank = int(os.environ.get('RANK', 0))
  world = int(os.environ.get('WORLD_SIZE', 1))
  local = int(os.environ.get('LOCAL_RANK', 0))
  if args.impl == 'reference':
    return run_reference(args, rank, world)

  torch.cuda.set_device(local)
  dev = torch.device('cuda', local)
  dist = None
  if world > 1:
    import torch.distributed as dist
    dist.init_process_group('nccl', device_id=dev)

  sys.path.insert(0, os.path.join(ROOT, 'oracle'))
  sys.path.insert(0, os.path.join(ROOT, 'tests'))
  from centertrack_b200 import synthetic as wt
  from helpers import make_model
  from centertrack_b200 import _lib as L
  from centertrack_b200.runner import StreamRunner

  B = args.batch
  opt, model, sd = make_model('coco_tracking')
  model = model.to(dev)
  runner = StreamRunner(model, B, H, W, K=K, precision=args.precision, device=dev)
  # synthetic inputs: 2 slots x B distinct frames (+ pre_hm); inputs alone are 2 x B x 4.2 MB and one step
  # streams ~0.3 GB of activations per frame, so nothing but the 40 MB of weights can live in the 126 MB L2
  img, pre, hm = wt.synthetic_inputs(2, H, W, seed=317 + rank)
  g = torch.Generator().manual_seed(rank)
  host_img = [(img[s:s + 1] + 0.05 * torch.randn(B, 3, H, W, generator=g)).pin_memory() for s in range(2)]
  host_hm = [hm[s:s + 1].expand(B, 1, H, W).contiguous().pin_memory() for s in range(2)]
  from centertrack_b200.runner import NS
  for s in range(NS):
    runner.load_device_inputs(host_img[s & 1].to(dev), host_hm[s & 1].to(dev), s)
  runner.warm()

  def barrier():
    if dist is not None:
      dist.barrier()
    torch.cuda.synchronize(dev)

  gathered = torch.empty((world * B,) + tuple(runner.rec.shape[1:]), device=dev) if world > 1 else None

  # ---------------- device-resident timing (value) ----------------
  def dev_step():
    runner.step_device()
    if gathered is not None:               # the one collective of the path: fixed-size result gather
      dist.all_gather_into_tensor(gathered, runner.rec)

  for _ in range(args.warmup):
    dev_step()
  barrier()
  sampler = ClockSampler(local)
  if rank == 0:
    sampler.start()
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  e0.record()
  for _ in range(args.steps):
    dev_step()
  e1.record()
  barrier()
  ms = e0.elapsed_time(e1)
  if dist is not None:
    t = torch.tensor([ms], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
  clocks = sampler.summary() if rank == 0 else None
  ms_per_step = ms / args.steps
  value = world * B * args.steps / (ms / 1000.0)

  # ---------------- end-to-end timing (host buffers in, host records out) ----------------
  for i in range(args.warmup):
    runner.step_host(host_img[i & 1], host_hm[i & 1])
  runner.fetch()
  barrier()
  t0 = time.perf_counter()
  for i in range(args.steps):
    runner.step_host(host_img[i & 1], host_hm[i & 1])
  rec_last = runner.fetch()
  barrier()
  e2e_s = time.perf_counter() - t0
  if dist is not None:
    t = torch.tensor([e2e_s], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_s = float(t.item())
  e2e = world * B * args.steps / e2e_s

  if rank != 0:
    if dist is not None:
      dist.destroy_process_group()
    return

  # ---------------- roofline of the dominant kernel (conv_tc_kernel), measured live ----------------
  # `traffic`: dram__bytes_read.sum + dram__bytes_write.sum of these launches from the committed ncu --set full capture
  # of one step at 32 frames/step (profiles/r01f_ncu_tc_raw.csv: 2157.5 + 381.1 MB over the 40 launches;
  # profiles/r01f_ncu_halo_raw.csv: 3058.0 + 1848.4 MB over the 34 launches), scaled by frames per step.
  NCU_DRAM_BYTES_PER_FRAME = {'conv_tc': 2538.59e6 / 32, 'conv_halo': 4906.37e6 / 32}
  eng = runner.eng
  stream = torch.cuda.current_stream()
  conv_ms, conv_flop, all_ms, halo_ms, halo_flop = 0.0, 0.0, 0.0, 0.0, 0.0
  reps = 3
  per_kind = {}
  for rep in range(reps):
    marks = [torch.cuda.Event(enable_timing=True) for _ in range(len(eng.ops) + 1)]
    marks[0].record()
    st = L.stream_ptr()
    for i, (kind, pl, name) in enumerate(eng.ops):
      eng._run_one(kind, pl, name, L.ptr(runner.img[0]), L.ptr(runner.img[1]), L.ptr(runner.hm[0]), st)
      marks[i + 1].record()
    torch.cuda.synchronize(dev)
    for i, (kind, pl, name) in enumerate(eng.ops):
      dt = marks[i].elapsed_time(marks[i + 1])
      all_ms += dt
      key = 'dcn' if (kind == 'conv' and pl.a_mode == L.CT_A_DCN) else kind
      per_kind[key] = per_kind.get(key, 0.0) + dt / reps
      fl = 2.0 * pl.B * pl.OH * pl.OW * pl.C_out * pl.KH * pl.KW * pl.C_in if kind == 'conv' else 0.0
      if kind == 'conv' and pl.epilogue_sum3:
        fl = 2.0 * pl.B * pl.OH * pl.OW * 16 * 49 * 7            # the three stems: 16 x (3+3+1) x 7x7 MACs/pixel
      if kind == 'conv' and pl.engine == L.CT_ENGINE_TCGEN05:
        conv_ms += dt
        conv_flop += fl
      if kind == 'conv' and pl.engine == L.CT_ENGINE_TCGEN05_HALO:
        halo_ms += dt
        halo_flop += fl
  conv_ms /= reps
  conv_flop /= reps
  halo_ms /= reps
  halo_flop /= reps
  all_ms /= reps
  # decode-only latency: the fused NMS + top-K + gather launch on the maps the last step left in HBM
  from centertrack_b200.decode import generic_decode
  dec_out = dict(eng.forward(runner.img[0], runner.img[1], runner.hm[0]))
  for _ in range(3):
    generic_decode(dec_out, K=K, records_out=runner.rec)
  d0, d1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  d0.record()
  for _ in range(10):
    generic_decode(dec_out, K=K, records_out=runner.rec)
  d1.record()
  torch.cuda.synchronize(dev)
  decode_us = d0.elapsed_time(d1) * 1000.0 / 10
  hm_bytes = float(dec_out['hm'].numel() * 4)
  peak_tf, peak_hbm, peak_src = _peaks()
  achieved = conv_flop / (conv_ms / 1000.0) / 1e12 if conv_ms > 0 else 0.0
  halo_tf = halo_flop / (halo_ms / 1000.0) / 1e12 if halo_ms > 0 else 0.0
  roofline = {'kernel': 'conv_tc_kernel (all %d gather-engine tcgen05 conv/DCN launches of one step)' %
              sum(1 for k, p, n in eng.ops if k == 'conv' and p.engine == L.CT_ENGINE_TCGEN05),
              'bound': 'tensor', 'achieved': achieved, 'peak': peak_tf, 'unit': 'TFLOP/s',
              'frac': achieved / peak_tf if peak_tf else None,
              'traffic': NCU_DRAM_BYTES_PER_FRAME['conv_tc'] * B if args.precision == 'bf16' else None,
              'traffic_unit': 'bytes per step (all launches of the kernel), ncu capture at 32 frames/step scaled',
              'peak_source': peak_src,
              'share_of_step': conv_ms / all_ms if all_ms else None,
              'second_kernel': {'kernel': 'conv_halo_kernel (%d launches)' % sum(1 for k, p, n in eng.ops if k == 'conv' and p.engine == L.CT_ENGINE_TCGEN05_HALO),
                                'achieved': halo_tf, 'frac': halo_tf / peak_tf if peak_tf else None,
                                'traffic': NCU_DRAM_BYTES_PER_FRAME['conv_halo'] * B if args.precision == 'bf16' else None,
                                'share_of_step': halo_ms / all_ms if all_ms else None},
              'eager_ms_by_kind': {k: round(v, 3) for k, v in per_kind.items()},
              'decode': {'us_per_launch': round(decode_us, 1), 'frames': B, 'bound': 'hbm',
                         'achieved': hm_bytes / (decode_us * 1e-6) / 1e9, 'peak': peak_hbm, 'unit': 'GB/s',
                         'frac': (hm_bytes / (decode_us * 1e-6) / 1e9) / peak_hbm if peak_hbm else None},
              'whole_step_tflops': GFLOP_PER_FRAME * B / ms_per_step}

  cpu = None
  if not args.no_cpu_baseline and world == 1:       # rank 0 at N=1 only
    _use_host_threads()
    _oracle_step(1)                                     # warm-up (oneDNN primitive caches)
    dt, thr, n = _oracle_step(10, budget_s=15.0)        # ~10 s of wall clock on the box's 64 host threads
    cpu = {'value': n / dt, 'unit': 'frames/s', 'cores': thr, 'kind': 'port',
           'sample': '%d frames 512x512 (oracle/ct_oracle.py, torch-CPU fp32)' % n}

  line = {'metric': METRIC, 'value': value, 'unit': 'frames/s', 'n_gpus': world, 'steps': args.steps,
          'warmup': args.warmup, 'ms_per_step': ms_per_step, 'higher_is_better': True, 'scaling': 'weak',
          'vs_baseline': None, 'dtype': args.precision, 'data': 'synthetic',
          'config': {'workload': 'DLA-34 coco_tracking 512x512 frame pairs + pre_hm, K=100 (BASELINE configs[1])',
                     'frames_per_step_per_gpu': B, 'global_batch': B * world,
                     'parallelism': 'stream-sharded replicas x%d, NCCL all_gather of records' % world,
                     'l2': 'no flush: per-step inputs %.0f MB in 3 rotating slots + ~%.1f GB of activations per '
                           'step exceed the 126 MB L2' % (2 * B * 4 * H * W * 4 / 1e6, 0.29 * B),
                     'cuda_graph': True},
          'e2e': {'value': e2e, 'unit': 'frames/s', 'h2d_bytes_per_step': runner.h2d_bytes_per_step,
                  'd2h_bytes_per_step': runner.d2h_bytes_per_step},
          'gpu_launches': int(runner.launches_per_step * args.steps),
          'clocks': clocks, 'roofline': roofline, 'cpu_baseline': cpu,
          'check': {'top_score_frame0': float(rec_last[0, 0, 0])}}
  _emit(line)
  if dist is not None:
    dist.destroy_process_group()


if __name__ == '__main__':
  main()
