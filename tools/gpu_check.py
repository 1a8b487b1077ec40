"""Developer bring-up script (GPU box): python tools/gpu_check.py <section>
sections: decode | conv_simt | conv_tc | dcn | net_fp32 | net_bf16 | time
Each section runs in its own process (tools/gpu_check.sh) so a trap in one kernel cannot mask others."""
import ctypes as C
import os
import sys
import time
import traceback

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'oracle'))
sys.path.insert(0, os.path.join(ROOT, 'tests'))

from centertrack_b200 import _lib as L   # noqa
import ct_oracle as co                    # noqa
from centertrack_b200 import synthetic as wt   # noqa

dev = torch.device('cuda')


def stat(name, got, ref, tol):
  got, ref = got.float().cpu(), ref.float().cpu()
  d = (got - ref).abs()
  err = d.max().item()
  rel = err / max(ref.abs().max().item(), 1e-12)
  ok = err <= tol * max(1.0, ref.abs().max().item())
  print('%-40s max %.3e rel %.3e mean %.3e p99.9 %.3e ref_max %.2e ref_std %.2e %s' % (
      name, err, rel, d.mean().item(), torch.quantile(d.flatten()[:4000000], 0.999).item(),
      ref.abs().max().item(), ref.std().item(), 'OK' if ok else 'FAIL'))
  return ok


# ------------------------------------------------------------------------------------------
from gpu_helpers import run_conv, conv_cases   # noqa


def sec_conv(engine, dtype, tol):
  g = torch.Generator().manual_seed(0)
  ok = True
  for (name, B, Cin, Cout, H, W, k, s, res, ld_pad, ch_off) in conv_cases():
    x = torch.randn(B, Cin, H, W, generator=g)
    w = torch.randn(Cout, Cin, k, k, generator=g) * (2.0 / (Cin * k * k)) ** 0.5
    b = torch.randn(Cout, generator=g) * 0.1
    OH, OW = (H + 2 * (k // 2) - k) // s + 1, (W + 2 * (k // 2) - k) // s + 1
    r = torch.randn(B, Cout, OH, OW, generator=g) if res else None
    if dtype == L.CT_BF16:
      xq, rq = x.bfloat16().float(), (r.bfloat16().float() if res else None)
      wq = w.bfloat16().float() if engine == L.CT_ENGINE_TCGEN05 else w
    else:
      xq, rq, wq = x, r, w
    ref = F.conv2d(xq, wq, b, s, k // 2)
    if res:
      ref = ref + rq
    ref = F.relu(ref)
    try:
      got = run_conv(engine, dtype, x.to(dev), w, b, s, True, r.to(dev) if res else None, ld_pad=ld_pad, ch_off=ch_off)
      ok &= stat(name, got, ref, tol)
    except Exception:
      traceback.print_exc()
      ok = False
  # head-style NCHW fp32 output with sigmoid, C_out=2 / 80
  for Cout, act in ((2, 0), (80, 1), (1, 2)):
    x = torch.randn(1, 256, 16, 24, generator=g)
    w = torch.randn(Cout, 256, 1, 1, generator=g) * 0.05
    b = torch.randn(Cout, generator=g)
    xq = x.bfloat16().float() if dtype == L.CT_BF16 else x
    wq = w.bfloat16().float() if engine == L.CT_ENGINE_TCGEN05 else w
    ref = F.conv2d(xq, wq, b)
    if act == 1:
      ref = torch.sigmoid(ref)
    if act == 2:
      ref = 1. / (torch.sigmoid(ref) + 1e-6) - 1.
    try:
      got = run_conv(engine, dtype, x.to(dev), w, b, 1, False, out_mode=L.CT_OUT_NCHW_F32, head_act=act)
      ok &= stat('1x1 256->%d NCHW f32 act%d' % (Cout, act), got, ref, tol * (10 if act == 2 else 1))
    except Exception:
      traceback.print_exc()
      ok = False
  return ok


def sec_halo(tol=6e-3):
  """CT_ENGINE_TCGEN05_HALO vs torch fp32 on bf16-rounded operands."""
  g = torch.Generator().manual_seed(11)
  E, D = L.CT_ENGINE_TCGEN05_HALO, L.CT_BF16
  ok = True
  cases = [('3x3 64->64 res', 2, 64, 64, 24, 40, 3, True, 0), ('3x3 16->16', 1, 16, 16, 40, 56, 3, False, 0),
           ('3x3 32->64', 1, 32, 64, 20, 28, 3, False, 0), ('1x1 64->32', 1, 64, 32, 16, 24, 1, False, 0),
           ('3x3 64->1024 nt128', 1, 64, 1024, 16, 24, 3, False, 128), ('3x3 48->16', 1, 48, 16, 33, 17, 3, False, 0),
           ('3x3 64->64 big', 4, 64, 64, 128, 128, 3, True, 0), ('3x3 128->128 2 chunks nt32', 2, 128, 128, 24, 40, 3, True, 32),
           ('1x1 256->128 4 chunks', 1, 256, 128, 16, 24, 1, False, 128), ('3x3 128->64 nt32', 1, 128, 64, 20, 28, 3, False, 32)]
  for (name, B, Cin, Cout, H, W, k, res, nt) in cases:
    x = torch.randn(B, Cin, H, W, generator=g)
    w = torch.randn(Cout, Cin, k, k, generator=g) * (2.0 / (Cin * k * k)) ** 0.5
    b = torch.randn(Cout, generator=g) * 0.1
    r = torch.randn(B, Cout, H, W, generator=g) if res else None
    ref = F.conv2d(x.bfloat16().float(), w.bfloat16().float(), b, 1, k // 2)
    if res:
      ref = ref + r.bfloat16().float()
    ref = F.relu(ref)
    try:
      got = run_conv(E, D, x.to(dev), w, b, 1, True, r.to(dev) if res else None, n_tile=nt)
      ok &= stat('halo ' + name, got, ref, tol)
    except Exception:
      traceback.print_exc()
      ok = False
  # DCN offset-style output
  x = torch.randn(1, 64, 24, 40, generator=g)
  w = torch.randn(27, 64, 3, 3, generator=g) * 0.03
  b = torch.randn(27, generator=g)
  ref = F.conv2d(x.bfloat16().float(), w.bfloat16().float(), b, 1, 1)
  ref[:, 18:] = torch.sigmoid(ref[:, 18:])
  try:
    got = run_conv(E, D, x.to(dev), w, b, 1, relu=False, out_mode=L.CT_OUT_NHWC_F32, sig_from=18, n_tile=32)
    ok &= stat('halo 3x3 64->27 f32 nhwc', got[:, :27], ref, 1e-4)
  except Exception:
    traceback.print_exc(); ok = False
  # heads-style NCHW output
  x = torch.randn(2, 64, 16, 24, generator=g)
  w = torch.randn(80, 64, 1, 1, generator=g) * 0.1
  b = torch.randn(80, generator=g)
  ref = torch.sigmoid(F.conv2d(x.bfloat16().float(), w.bfloat16().float(), b))
  try:
    got = run_conv(E, D, x.to(dev), w, b, 1, relu=False, out_mode=L.CT_OUT_NCHW_F32, head_act=1, n_tile=80)
    ok &= stat('halo 1x1 64->80 nchw sigmoid', got, ref, 1e-4)
  except Exception:
    traceback.print_exc(); ok = False
  x = torch.randn(2, 256, 16, 24, generator=g)
  w = torch.randn(80, 256, 1, 1, generator=g) * 0.05
  b = torch.randn(80, generator=g)
  ref = torch.sigmoid(F.conv2d(x.bfloat16().float(), w.bfloat16().float(), b))
  try:
    got = run_conv(E, D, x.to(dev), w, b, 1, relu=False, out_mode=L.CT_OUT_NCHW_F32, head_act=1, n_tile=80)
    ok &= stat('halo 1x1 256->80 nchw sigmoid (head)', got, ref, 1e-4)
  except Exception:
    traceback.print_exc(); ok = False
  # stem: 7x7, C_in = 8 (img3, pre3, hm1, 0), block-diagonal 48 outputs, relu per group then sum
  for mask in (7, 1, 3):
    img, pre, hm = torch.randn(2, 3, 40, 56, generator=g), torch.randn(2, 3, 40, 56, generator=g), torch.rand(2, 1, 40, 56, generator=g)
    ws = [torch.randn(16, c, 7, 7, generator=g) * 0.1 for c in (3, 3, 1)]
    sh = torch.randn(48, generator=g) * 0.2
    w48 = torch.zeros(48, 8, 7, 7)
    w48[0:16, 0:3], w48[16:32, 3:6], w48[32:48, 6:7] = ws[0], ws[1], ws[2]
    x8 = torch.cat([img, pre if mask & 2 else torch.zeros_like(pre), hm if mask & 4 else torch.zeros_like(hm),
                    torch.zeros(2, 1, 40, 56)], 1)
    ref = 0
    for gi, (t, wgt) in enumerate(zip((img, pre, hm), ws)):
      if (mask >> gi) & 1:
        ref = ref + F.relu(F.conv2d(t.bfloat16().float(), wgt.bfloat16().float(), sh[16 * gi:16 * gi + 16], 1, 3))
    try:
      got = run_conv(E, D, x8.to(dev), w48, sh, 1, relu=False, n_tile=48, sum3=mask)
      ok &= stat('halo stem 7x7 8->48 sum3 mask%d' % mask, got, ref, tol)
    except Exception:
      traceback.print_exc(); ok = False
  return ok


def sec_dcn(engine, dtype, tol):
  g = torch.Generator().manual_seed(1)
  ok = True
  for (B, Cin, Cout, H, W) in ((1, 64, 64, 24, 40), (2, 128, 64, 16, 16), (1, 256, 256, 8, 12), (1, 512, 256, 4, 6)):
    x = torch.randn(B, Cin, H, W, generator=g)
    w = torch.randn(Cout, Cin, 3, 3, generator=g) * (2.0 / (Cin * 9)) ** 0.5
    b = torch.randn(Cout, generator=g) * 0.1
    wo = torch.randn(27, Cin, 3, 3, generator=g) * 0.02
    bo = torch.randn(27, generator=g) * 0.3
    xq = x.bfloat16().float() if dtype == L.CT_BF16 else x
    tc = engine == L.CT_ENGINE_TCGEN05
    wq, woq = (w.bfloat16().float(), wo.bfloat16().float()) if tc else (w, wo)
    om_ref = F.conv2d(xq, woq, bo, 1, 1)
    try:
      om = run_conv(engine, dtype, x.to(dev), wo, bo, 1, relu=False, out_mode=L.CT_OUT_NHWC_F32, sig_from=18, n_tile=32)
      om_ref_s = om_ref.clone()
      om_ref_s[:, 18:] = torch.sigmoid(om_ref[:, 18:27])
      ok &= stat('DCN offset conv %d->27 %dx%d' % (Cin, H, W), om[:, :27], om_ref_s, tol)
      # main: feed the DEVICE offsets to both sides so sampling positions are identical
      om_dev = om.permute(0, 2, 3, 1).contiguous()            # [B,H,W,32] fp32
      om_cpu = om.cpu()
      cols = co.dcn_sample_columns(xq, om_cpu[:, :18], om_cpu[:, 18:27])
      if tc:
        cols = cols.bfloat16().float()
      ref = torch.einsum('ok,bkp->bop', wq.reshape(Cout, Cin * 9), cols.reshape(B, Cin * 9, H * W)).view(B, Cout, H, W)
      ref = F.relu(ref + b.view(1, -1, 1, 1))
      got = run_conv(engine, dtype, x.to(dev), w, b, 1, relu=True, a_mode=L.CT_A_DCN, om=om_dev)
      ok &= stat('DCN main %d->%d %dx%d' % (Cin, Cout, H, W), got, ref, tol)
    except Exception:
      traceback.print_exc()
      ok = False
  return ok


def sec_decode():
  from centertrack_b200.decode import generic_decode
  ok = True
  rng = np.random.RandomState(7)
  cases = [(1, 80, 128, 128, 100, 'coco'), (2, 1, 136, 240, 100, 'mot'), (1, 10, 112, 200, 100, 'ddd'),
           (1, 1, 128, 128, 100, 'pose'), (3, 5, 32, 32, 40, 'ties'), (1, 3, 8, 8, 64, 'tiny')]
  for (B, Cc, H, W, K, kind) in cases:
    hm = 1. / (1. + np.exp(-(2 * rng.randn(B, Cc, H, W) - 4.6)))
    hm = hm.astype(np.float32)
    if kind == 'ties':
      hm = np.round(hm * 50) / 50        # massive value ties + plateaus
      hm = hm.astype(np.float32)
    out = {'hm': hm, 'reg': rng.randn(B, 2, H, W).astype(np.float32), 'wh': (rng.randn(B, 2, H, W) * 5).astype(np.float32),
           'tracking': rng.randn(B, 2, H, W).astype(np.float32)}
    if kind == 'ddd':
      out.update({'dep': rng.rand(B, 1, H, W).astype(np.float32) * 50, 'rot': rng.randn(B, 8, H, W).astype(np.float32),
                  'dim': rng.randn(B, 3, H, W).astype(np.float32), 'amodel_offset': rng.randn(B, 2, H, W).astype(np.float32)})
    if kind == 'pose':
      hp = 1. / (1. + np.exp(-(2 * rng.randn(B, 17, H, W) - 3.0)))
      out.update({'hps': (rng.randn(B, 34, H, W) * 6).astype(np.float32), 'hm_hp': hp.astype(np.float32),
                  'hp_offset': rng.rand(B, 2, H, W).astype(np.float32)})
    if kind == 'mot':
      out['ltrb_amodal'] = (rng.randn(B, 4, H, W) * 8).astype(np.float32)
    ref = co.generic_decode(out, K)
    dout = {k: torch.from_numpy(v).to(dev) for k, v in out.items()}
    t0 = time.time()
    got = generic_decode(dout, K=K)
    torch.cuda.synchronize()
    inds = got.inds.cpu().numpy()
    exact = np.array_equal(inds, ref['_inds'].astype(np.int32))
    print('[decode %s] B%d C%d %dx%d K%d  top-K indices bit-exact: %s' % (kind, B, Cc, H, W, K, exact))
    ok &= exact
    for k in ref:
      if k.startswith('_'):
        continue
      g = got[k].cpu().numpy().reshape(ref[k].shape)
      tol = 0 if k in ('scores', 'clses', 'xs', 'ys', 'cts', 'bboxes', 'tracking', 'dep', 'rot', 'dim', 'amodel_offset',
                       'bboxes_amodal') else 1e-5
      e = np.abs(g - ref[k]).max()
      good = e <= tol
      if not good:
        print('   key %-14s max err %.3e  FAIL' % (k, e))
      ok &= bool(good)
    # timing
    for _ in range(3):
      generic_decode(dout, K=K)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
      generic_decode(dout, K=K)
    e1.record()
    torch.cuda.synchronize()
    print('   decode latency (incl. python launch overhead): %.1f us' % (e0.elapsed_time(e1) * 1000 / 20))
  return ok


def build_models(cfg_heads, H, W, B=1, seed=317):
  from centertrack_b200.model import create_model
  from centertrack_b200.opts import opts
  task = {'coco_tracking': 'tracking', 'nuscenes_ddd': 'tracking,ddd', 'coco_pose': 'tracking,multi_pose'}[cfg_heads]
  opt = opts().init([task, '--pre_hm'])
  m = create_model(opt.arch, opt.heads, opt.head_conv, opt=opt)
  sd = wt.make_state_dict(m.state_dict(), seed)
  m.load_state_dict(sd)
  return opt, m, sd


def sec_net(precision, tol, cfg='coco_tracking', H=64, W=96, emulate=False):
  opt, m, sd = build_models(cfg, H, W)
  img, pre, hm = wt.synthetic_inputs(1, H, W)
  trace = {}
  ref = co.DLA34Oracle(sd, opt.heads, emulate_bf16=emulate).forward(img, pre, hm, trace=trace)
  m = m.to(dev)
  eng = m.engine_for(1, H, W, dev, precision)
  out = eng.forward(img.to(dev), pre.to(dev), hm.to(dev))
  torch.cuda.synchronize()
  ok = True
  for name in ['stem', 'base.level0', 'base.level1', 'base.level2', 'base.level3', 'base.level4', 'base.level5',
               'dla_up.ida_0.proj_1', 'dla_up.ida_0.node_1', 'dla_up.ida_1.node_2', 'dla_up.ida_2.node_3',
               'ida_up.node_1', 'feat']:
    if name in trace and name in eng.named:
      ok &= stat('[%s] %s' % (precision, name), eng.stage(name), trace[name], tol)
  for h in opt.heads:
    ok &= stat('[%s] head %s' % (precision, h), out[h], ref[h], tol)
  # graph replay == eager
  eng.in_img.copy_(img.to(dev)); eng.in_pre.copy_(pre.to(dev)); eng.in_hm.copy_(hm.to(dev))
  eager = {h: out[h].clone() for h in out}
  rep = eng.replay()
  torch.cuda.synchronize()
  same = all(torch.equal(eager[h], rep[h]) for h in eager)
  print('graph replay bit-identical to eager:', same)
  return ok and same


def sec_time():
  from centertrack_b200.decode import generic_decode
  for precision in ('bf16', 'fp32'):
    for B in ((1, 8, 16) if precision == 'bf16' else (1,)):
      H = W = 512
      opt, m, sd = build_models('coco_tracking', H, W)
      m = m.to(dev)
      eng = m.engine_for(B, H, W, dev, precision)
      eng.set_fused_activations(True)
      img, pre, hm = wt.synthetic_inputs(1, H, W)
      eng.in_img.copy_(img.to(dev).expand(B, -1, -1, -1)); eng.in_pre.copy_(pre.to(dev).expand(B, -1, -1, -1))
      eng.in_hm.copy_(hm.to(dev).expand(B, -1, -1, -1))
      eng.replay()
      torch.cuda.synchronize()
      e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
      n = 20 if precision == 'bf16' else 3
      e0.record()
      for _ in range(n):
        eng.replay()
      e1.record()
      torch.cuda.synchronize()
      ms = e0.elapsed_time(e1) / n
      print('[time] %s B=%d 512x512: %.3f ms/step  %.1f frames/s  (%d launches/step)' % (precision, B, ms, B * 1000 / ms, eng.n_launches))
      if precision == 'bf16' and B == 1:
        # per-op eager timing
        torch.cuda.synchronize()
        times = []
        for kind, pl, name in eng.ops:
          pass
      del eng, m
      torch.cuda.empty_cache()
  return True


if __name__ == '__main__':
  sec = sys.argv[1]
  t0 = time.time()
  print('==== section %s on %s' % (sec, torch.cuda.get_device_name(0)))
  if sec == 'decode':
    ok = sec_decode()
  elif sec == 'conv_simt':
    ok = sec_conv(L.CT_ENGINE_SIMT, L.CT_F32, 2e-5) & sec_conv(L.CT_ENGINE_SIMT, L.CT_BF16, 6e-3)
  elif sec == 'conv_tc':
    ok = sec_conv(L.CT_ENGINE_TCGEN05, L.CT_BF16, 6e-3)
  elif sec == 'conv_halo':
    ok = sec_halo()
  elif sec == 'dcn':
    ok = sec_dcn(L.CT_ENGINE_SIMT, L.CT_F32, 5e-5) & sec_dcn(L.CT_ENGINE_TCGEN05, L.CT_BF16, 8e-3)
  elif sec == 'net_fp32':
    ok = sec_net('fp32', 1e-3)
  elif sec == 'net_bf16':
    ok = sec_net('bf16', 6e-2)
  elif sec == 'net_bf16_emu':
    ok = sec_net('bf16', 2e-2, emulate=True)
  elif sec == 'time':
    ok = sec_time()
  print('==== section %s %s (%.1fs)' % (sec, 'PASSED' if ok else 'FAILED', time.time() - t0))
  sys.exit(0 if ok else 1)
