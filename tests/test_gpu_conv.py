"""-m gpu: the conv / DCN engines through ct_conv_forward (C ABI) vs torch fp32 on CPU.
Tolerances: SIMT fp32 <= 2e-5 x scale (fp32 summation order); bf16 engines are compared against the
fp32 result on bf16-ROUNDED operands, <= 6e-3 x scale = one bf16 output rounding (2^-8) + slack; the bf16x3
tensor-core engine against the plain fp32 result, <= 6e-5 x scale."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

import ct_oracle as co
from centertrack_b200 import _lib as L

pytestmark = pytest.mark.gpu

ENGINES = [('simt_f32', L.CT_ENGINE_SIMT, L.CT_F32, 2e-5), ('simt_bf16', L.CT_ENGINE_SIMT, L.CT_BF16, 6e-3),
           ('tcgen05', L.CT_ENGINE_TCGEN05, L.CT_BF16, 6e-3),
           # bf16 hi/lo split operands on fp32 activations, three MMAs per term: ~2^-16 relative per product
           ('tcgen05_x3', L.CT_ENGINE_TCGEN05_X3, L.CT_F32, 6e-5)]


def _close(got, ref, tol):
  err = (got.float().cpu() - ref).abs().max().item()
  assert err <= tol * max(1.0, ref.abs().max().item()), 'max err %.3e (ref max %.3e)' % (err, ref.abs().max().item())


def _case_list():
  from gpu_helpers import conv_cases
  return conv_cases()


@pytest.mark.parametrize('eng', ENGINES, ids=[e[0] for e in ENGINES])
@pytest.mark.parametrize('case', range(7))
def test_conv_bn_residual_relu(eng, case):
  from gpu_helpers import run_conv
  _, engine, dtype, tol = eng
  name, B, Cin, Cout, H, W, k, s, res, ld_pad, ch_off = _case_list()[case]
  g = torch.Generator().manual_seed(case)
  x = torch.randn(B, Cin, H, W, generator=g)
  w = torch.randn(Cout, Cin, k, k, generator=g) * (2.0 / (Cin * k * k)) ** 0.5
  b = torch.randn(Cout, generator=g) * 0.1
  OH, OW = (H + 2 * (k // 2) - k) // s + 1, (W + 2 * (k // 2) - k) // s + 1
  r = torch.randn(B, Cout, OH, OW, generator=g) if res else None
  bf = dtype == L.CT_BF16
  xq = x.bfloat16().float() if bf else x
  rq = (r.bfloat16().float() if bf else r) if res else None
  wq = w.bfloat16().float() if engine == L.CT_ENGINE_TCGEN05 else w
  ref = F.conv2d(xq, wq, b, s, k // 2)
  ref = F.relu(ref + rq if res else ref)
  got = run_conv(engine, dtype, x.cuda(), w, b, s, True, r.cuda() if res else None, ld_pad=ld_pad, ch_off=ch_off)
  _close(got, ref, tol)


@pytest.mark.parametrize('eng', ENGINES, ids=[e[0] for e in ENGINES])
@pytest.mark.parametrize('cout_act', [(2, 0), (80, 1), (1, 2), (17, 1)])
def test_head_1x1_writes_reference_layout_with_fused_activation(eng, cout_act):
  from gpu_helpers import run_conv
  _, engine, dtype, tol = eng
  Cout, act = cout_act
  g = torch.Generator().manual_seed(Cout)
  x = torch.randn(2, 256, 16, 24, generator=g)
  w = torch.randn(Cout, 256, 1, 1, generator=g) * 0.05
  b = torch.randn(Cout, generator=g)
  xq = x.bfloat16().float() if dtype == L.CT_BF16 else x
  wq = w.bfloat16().float() if engine == L.CT_ENGINE_TCGEN05 else w
  ref = F.conv2d(xq, wq, b)
  if act == 1:
    ref = torch.sigmoid(ref)
  if act == 2:
    ref = 1. / (torch.sigmoid(ref) + 1e-6) - 1.
  got = run_conv(engine, dtype, x.cuda(), w, b, 1, False, out_mode=L.CT_OUT_NCHW_F32, head_act=act)
  _close(got, ref, tol * (10 if act == 2 else 1))


DCN_ENGINES = [ENGINES[0], ENGINES[2], ENGINES[2] + ('win',), ENGINES[3]]


@pytest.mark.parametrize('eng', DCN_ENGINES, ids=['simt_f32', 'tcgen05', 'tcgen05_window', 'tcgen05_x3'])
@pytest.mark.parametrize('shape', [(1, 64, 64, 24, 40), (2, 128, 64, 16, 16), (1, 256, 256, 8, 12),
                                   (1, 512, 256, 4, 6), (1, 64, 64, 5, 7), (2, 64, 64, 24, 32), (3, 64, 128, 40, 56),
                                   (4, 64, 64, 64, 96), (4, 128, 64, 64, 96)])     # > 148 patches: several tiles per persistent CTA
def test_dcn_v2(eng, shape):
  """Offset/mask conv + modulated deformable conv; large offsets push samples across and beyond the
  border (zero padding, partial bilinear weights) and, for the shared-memory-window variant (CT_A_DCN_WIN), beyond
  the staged window (global fall-back path) -- ragged 8x16 patches included."""
  from gpu_helpers import run_conv
  _, engine, dtype, tol = eng[:4]
  dcn_mode = L.CT_A_DCN_WIN if len(eng) > 4 else L.CT_A_DCN
  B, Cin, Cout, H, W = shape
  g = torch.Generator().manual_seed(H * W)
  x = torch.randn(B, Cin, H, W, generator=g)
  w = torch.randn(Cout, Cin, 3, 3, generator=g) * (2.0 / (Cin * 9)) ** 0.5
  b = torch.randn(Cout, generator=g) * 0.1
  wo = torch.randn(27, Cin, 3, 3, generator=g) * (0.6 / (Cin * 9) ** 0.5)
  bo = torch.randn(27, generator=g) * 1.5
  tc = engine == L.CT_ENGINE_TCGEN05
  xq = x.bfloat16().float() if tc else x
  wq, woq = (w.bfloat16().float(), wo.bfloat16().float()) if tc else (w, wo)
  om_ref = F.conv2d(xq, woq, bo, 1, 1)
  om_ref[:, 18:] = torch.sigmoid(om_ref[:, 18:27])
  om = run_conv(engine, dtype, x.cuda(), wo, bo, 1, relu=False, out_mode=L.CT_OUT_NHWC_F32, sig_from=18, n_tile=32)
  _close(om[:, :27], om_ref, (6e-5 if engine == L.CT_ENGINE_TCGEN05_X3 else 2e-5) if not tc else 1e-4)
  # feed the DEVICE offsets to both sides so the sampling positions are identical
  om_dev = om.permute(0, 2, 3, 1).contiguous()
  omc = om.cpu()
  assert float(omc[:, :18].abs().max()) > 2.0          # the case really leaves the 3x3 window
  cols = co.dcn_sample_columns(xq, omc[:, :18], omc[:, 18:27], bf16_blend=tc)      # both bf16 samplers blend in packed bf16
  if tc:
    cols = cols.bfloat16().float()
  ref = torch.einsum('ok,bkp->bop', wq.reshape(Cout, Cin * 9), cols.reshape(B, Cin * 9, H * W)).view(B, Cout, H, W)
  ref = F.relu(ref + b.view(1, -1, 1, 1))
  got = run_conv(engine, dtype, x.cuda(), w, b, 1, relu=True, a_mode=dcn_mode, om=om_dev)
  _close(got, ref, (1e-4 if engine == L.CT_ENGINE_TCGEN05_X3 else 5e-5) if not tc else 8e-3)


HALO_CASES = [('3x3 64->64 +res', 2, 64, 64, 24, 40, 3, True, 0), ('3x3 16->16', 1, 16, 16, 40, 56, 3, False, 0),
              ('3x3 32->64 ragged tile', 1, 32, 64, 20, 28, 3, False, 0), ('1x1 64->32', 1, 64, 32, 16, 24, 1, False, 0),
              ('3x3 64->1024 8 n-tiles', 1, 64, 1024, 16, 24, 3, False, 128), ('3x3 48->16 odd size', 1, 48, 16, 33, 17, 3, False, 0),
              ('3x3 64->64 many tiles per CTA', 4, 64, 64, 128, 128, 3, True, 0),
              ('3x3 128->128 two 64-ch chunks', 2, 128, 128, 24, 40, 3, True, 32),
              ('1x1 256->128 four chunks', 1, 256, 128, 16, 24, 1, False, 128), ('3x3 128->64', 1, 128, 64, 20, 28, 3, False, 32)]


@pytest.mark.parametrize('case', range(len(HALO_CASES)), ids=[c[0] for c in HALO_CASES])
def test_halo_engine_conv(case):
  """CT_ENGINE_TCGEN05_HALO (TMA halo tile, taps by descriptor shift, persistent CTAs) vs torch fp32 on
  bf16-rounded operands; covers multi-tile persistence, n-tiling, image borders and ragged tiles."""
  from gpu_helpers import run_conv
  name, B, Cin, Cout, H, W, k, res, nt = HALO_CASES[case]
  g = torch.Generator().manual_seed(100 + case)
  x = torch.randn(B, Cin, H, W, generator=g)
  w = torch.randn(Cout, Cin, k, k, generator=g) * (2.0 / (Cin * k * k)) ** 0.5
  b = torch.randn(Cout, generator=g) * 0.1
  r = torch.randn(B, Cout, H, W, generator=g) if res else None
  ref = F.conv2d(x.bfloat16().float(), w.bfloat16().float(), b, 1, k // 2)
  ref = F.relu(ref + r.bfloat16().float() if res else ref)
  got = run_conv(L.CT_ENGINE_TCGEN05_HALO, L.CT_BF16, x.cuda(), w, b, 1, True, r.cuda() if res else None, n_tile=nt)
  _close(got, ref, 6e-3)


def test_halo_engine_fp32_outputs():
  from gpu_helpers import run_conv
  g = torch.Generator().manual_seed(21)
  x = torch.randn(1, 64, 24, 40, generator=g)
  w = torch.randn(27, 64, 3, 3, generator=g) * 0.03
  b = torch.randn(27, generator=g)
  ref = F.conv2d(x.bfloat16().float(), w.bfloat16().float(), b, 1, 1)
  ref[:, 18:] = torch.sigmoid(ref[:, 18:])
  got = run_conv(L.CT_ENGINE_TCGEN05_HALO, L.CT_BF16, x.cuda(), w, b, 1, relu=False, out_mode=L.CT_OUT_NHWC_F32,
                 sig_from=18, n_tile=32)
  _close(got[:, :27], ref, 1e-4)
  x = torch.randn(2, 64, 16, 24, generator=g)
  w = torch.randn(80, 64, 1, 1, generator=g) * 0.1
  b = torch.randn(80, generator=g)
  ref = torch.sigmoid(F.conv2d(x.bfloat16().float(), w.bfloat16().float(), b))
  got = run_conv(L.CT_ENGINE_TCGEN05_HALO, L.CT_BF16, x.cuda(), w, b, 1, relu=False, out_mode=L.CT_OUT_NCHW_F32,
                 head_act=1, n_tile=80)
  _close(got, ref, 1e-4)
  # 32 x 4 pixel tiles (W % 32 == 0): the 1x1 heads at full resolution, 256 input channels in four chunks
  x = torch.randn(2, 256, 8, 64, generator=g)
  w = torch.randn(80, 256, 1, 1, generator=g) * 0.05
  ref = torch.sigmoid(F.conv2d(x.bfloat16().float(), w.bfloat16().float(), b))
  got = run_conv(L.CT_ENGINE_TCGEN05_HALO, L.CT_BF16, x.cuda(), w, b, 1, relu=False, out_mode=L.CT_OUT_NCHW_F32,
                 head_act=1, n_tile=80)
  _close(got, ref, 2e-4)
  w2 = torch.randn(2, 256, 1, 1, generator=g) * 0.05
  ref = F.conv2d(x.bfloat16().float(), w2.bfloat16().float(), b[:2])
  got = run_conv(L.CT_ENGINE_TCGEN05_HALO, L.CT_BF16, x.cuda(), w2, b[:2], 1, relu=False, out_mode=L.CT_OUT_NCHW_F32,
                 n_tile=16)
  _close(got, ref, 2e-3)


@pytest.mark.parametrize('mask', [7, 1, 3])
def test_halo_engine_tensor_core_stem(mask):
  """7x7, C_in = 8 = (img3, pre3, hm1, 0), two taps per K=16 MMA, block-diagonal 48 outputs; the epilogue
  applies ReLU per stem and sums the PRESENT stems (mask bit g) -- dla.py:305-311."""
  from gpu_helpers import run_conv
  g = torch.Generator().manual_seed(mask)
  img, pre, hm = torch.randn(2, 3, 40, 56, generator=g), torch.randn(2, 3, 40, 56, generator=g), torch.rand(2, 1, 40, 56, generator=g)
  ws = [torch.randn(16, c, 7, 7, generator=g) * 0.1 for c in (3, 3, 1)]
  sh = torch.randn(48, generator=g) * 0.2
  w48 = torch.zeros(48, 8, 7, 7)
  w48[0:16, 0:3], w48[16:32, 3:6], w48[32:48, 6:7] = ws[0], ws[1], ws[2]
  x8 = torch.cat([img, pre, hm, torch.zeros(2, 1, 40, 56)], 1)
  ref = 0
  for gi, (t, wgt) in enumerate(zip((img, pre, hm), ws)):
    if (mask >> gi) & 1:
      ref = ref + F.relu(F.conv2d(t.bfloat16().float(), wgt.bfloat16().float(), sh[16 * gi:16 * gi + 16], 1, 3))
  got = run_conv(L.CT_ENGINE_TCGEN05_HALO, L.CT_BF16, x8.cuda(), w48, sh, 1, relu=False, n_tile=48, sum3=mask)
  _close(got, ref, 6e-3)


def test_pack_stem_input():
  import ctypes as C
  g = torch.Generator().manual_seed(4)
  img, pre, hm = torch.randn(2, 3, 9, 13, generator=g).cuda(), torch.randn(2, 3, 9, 13, generator=g).cuda(), torch.rand(2, 1, 9, 13, generator=g).cuda()
  out = torch.empty(2, 9, 13, 8, dtype=torch.bfloat16, device='cuda')
  L.check(L.lib().ct_pack_stem_input(L.ptr(img), L.ptr(pre), L.ptr(None), L.ptr(out), 2, 9, 13, L.stream_ptr()))
  ref = torch.cat([img, pre, torch.zeros_like(hm), torch.zeros_like(hm)], 1).permute(0, 2, 3, 1).bfloat16()
  assert torch.equal(out, ref)


def test_halo_even_kernel_and_space_to_depth_output():
  """The two pieces of level1-as-a-2x2-convolution (engine.py): a 2x2 stride-1 conv padded on the top / left only, and a
  3x3 conv whose NHWC output is written space-to-depth; then the composition against the 3x3 stride-2 conv itself."""
  from gpu_helpers import run_conv
  g = torch.Generator().manual_seed(21)
  B, H, W = 2, 48, 80
  x = torch.randn(B, 64, H, W, generator=g)
  w = torch.randn(32, 64, 2, 2, generator=g) * 0.08
  b = torch.randn(32, generator=g) * 0.1
  got = run_conv(L.CT_ENGINE_TCGEN05_HALO, L.CT_BF16, x.cuda(), w, b, 1, True, n_tile=32).cpu()
  xb, wb = x.bfloat16().float(), w.bfloat16().float()
  ref = F.relu(F.conv2d(F.pad(xb, (1, 0, 1, 0)), wb, b))
  assert got.shape == ref.shape
  assert (got - ref).abs().max() < 2e-2 * max(1.0, float(ref.abs().max()))

  x0 = torch.randn(B, 16, H, W, generator=g)
  w0 = torch.randn(16, 16, 3, 3, generator=g) * 0.15
  b0 = torch.randn(16, generator=g) * 0.1
  plain = run_conv(L.CT_ENGINE_TCGEN05_HALO, L.CT_BF16, x0.cuda(), w0, b0, 1, True, n_tile=16).cpu()
  s2d = run_conv(L.CT_ENGINE_TCGEN05_HALO, L.CT_BF16, x0.cuda(), w0, b0, 1, True, n_tile=16,
                 out_mode=L.CT_OUT_NHWC_S2D).cpu()
  assert torch.equal(plain, s2d)                      # same values, only the layout differs

  # the stem's sum-of-three epilogue writing space-to-depth (what level0-on-the-s2d-grid reads)
  x8 = torch.randn(B, 8, H, W, generator=g)
  w48 = torch.randn(48, 8, 7, 7, generator=g) * 0.05
  b48 = torch.randn(48, generator=g) * 0.1
  st_plain = run_conv(L.CT_ENGINE_TCGEN05_HALO, L.CT_BF16, x8.cuda(), w48, b48, 1, False, n_tile=48, sum3=7).cpu()
  st_s2d = run_conv(L.CT_ENGINE_TCGEN05_HALO, L.CT_BF16, x8.cuda(), w48, b48, 1, False, n_tile=48, sum3=7,
                    out_mode=L.CT_OUT_NHWC_S2D).cpu()
  assert st_plain.shape == (B, 16, H, W) and torch.equal(st_plain, st_s2d)

  # level2's inputs: 2x2 max-pool of a space-to-depth tensor, and the 3x3 stride-2 32 -> 64 conv as a 2x2 over 128 channels
  import ctypes as C
  y = torch.randn(B, 32, H, W, generator=g).bfloat16()
  ys = y.reshape(B, 32, H // 2, 2, W // 2, 2).permute(0, 2, 4, 3, 5, 1).reshape(B, H // 2, W // 2, 128).contiguous().cuda()
  po = torch.empty(B, H // 2, W // 2, 32, dtype=torch.bfloat16, device='cuda')
  L.check(L.lib().ct_maxpool2_s2d(L.ptr(ys), L.ptr(po), L.CT_BF16, B, H // 2, W // 2, 32, 128, 32, L.stream_ptr()))
  assert torch.equal(po.permute(0, 3, 1, 2).float().cpu(), F.max_pool2d(y.float(), 2, 2))
  from centertrack_b200.engine import s2d_weights_3x3_s2
  w2 = torch.randn(64, 32, 3, 3, generator=g) * 0.08
  b2 = torch.randn(64, generator=g) * 0.1
  via2 = run_conv(L.CT_ENGINE_TCGEN05_HALO, L.CT_BF16, ys.permute(0, 3, 1, 2).float(), s2d_weights_3x3_s2(w2), b2, 1, True,
                  n_tile=64).cpu()
  ref2 = F.relu(F.conv2d(y.float(), w2.bfloat16().float(), b2, 2, 1))
  assert (via2 - ref2).abs().max() < 2e-2 * max(1.0, float(ref2.abs().max()))

  # composition: 3x3 stride-2 16 -> 32 == 2x2 stride-1 over the space-to-depth view with the regrouped weights
  w1 = torch.randn(32, 16, 3, 3, generator=g) * 0.1
  b1 = torch.randn(32, generator=g) * 0.1
  w1s = torch.zeros(32, 64, 2, 2)
  tap = ((0, 1), (1, 0), (1, 1))
  for ky in range(3):
    for kx in range(3):
      (ty, sy), (tx, sx) = tap[ky], tap[kx]
      w1s[:, (sy * 2 + sx) * 16:(sy * 2 + sx) * 16 + 16, ty, tx] = w1[:, :, ky, kx]
  xs = plain.reshape(B, 16, H // 2, 2, W // 2, 2).permute(0, 3, 5, 1, 2, 4).reshape(B, 64, H // 2, W // 2)
  via = run_conv(L.CT_ENGINE_TCGEN05_HALO, L.CT_BF16, xs.cuda(), w1s, b1, 1, True, n_tile=32).cpu()
  direct = run_conv(L.CT_ENGINE_TCGEN05, L.CT_BF16, plain.cuda(), w1, b1, 2, True).cpu()
  ref1 = F.relu(F.conv2d(plain, w1.bfloat16().float(), b1, 2, 1))
  assert (via - ref1).abs().max() < 2e-2 * max(1.0, float(ref1.abs().max()))
  assert (via - direct).abs().max() < 2e-2 * max(1.0, float(ref1.abs().max()))


@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16])
def test_maxpool_and_upsample_add(dtype):
  lib = L.lib()
  import ctypes as C
  ct = L.CT_F32 if dtype == torch.float32 else L.CT_BF16
  g = torch.Generator().manual_seed(3)
  x = torch.randn(2, 32, 12, 20, generator=g).to(dtype)
  xn = x.permute(0, 2, 3, 1).contiguous().cuda()
  out = torch.zeros(2, 6, 10, 48, dtype=dtype, device='cuda')          # write into a slice of a wider buffer
  L.check(lib.ct_maxpool2(L.ptr(xn), C.c_void_p(out.data_ptr() + 8 * out.element_size()), ct, 2, 12, 20, 32, 32, 48,
                          L.stream_ptr()))
  ref = F.max_pool2d(x.float(), 2, 2)
  assert torch.equal(out[..., 8:40].permute(0, 3, 1, 2).float().cpu(), ref)
  assert float(out[..., :8].abs().max()) == 0 and float(out[..., 40:].abs().max()) == 0
  for f in (2, 4, 8):
    w = torch.rand(32, 1, 2 * f, 2 * f, generator=g)
    skip = torch.randn(2, 32, 12 * f, 20 * f, generator=g).to(dtype)
    o = torch.empty(2, 12 * f, 20 * f, 32, dtype=dtype, device='cuda')
    wd = w.reshape(32, 2 * f, 2 * f).permute(1, 2, 0).contiguous().cuda()     # channel-last
    sk = skip.permute(0, 2, 3, 1).contiguous().cuda()
    L.check(lib.ct_upsample_add(L.ptr(xn), L.ptr(sk), L.ptr(wd), L.ptr(o), ct, 2, 12, 20, 32, f, 32, 32, 32,
                                L.stream_ptr()))
    ref = F.conv_transpose2d(x.float(), w, None, stride=f, padding=f // 2, groups=32) + skip.float()
    tol = 1e-5 if dtype == torch.float32 else 2e-2
    assert (o.permute(0, 3, 1, 2).float().cpu() - ref).abs().max() < tol


def test_stem_three_inputs_relu_before_sum():
  lib = L.lib()
  g = torch.Generator().manual_seed(9)
  B, H, W = 2, 40, 56
  img, pre, hm = torch.randn(B, 3, H, W, generator=g), torch.randn(B, 3, H, W, generator=g), torch.rand(B, 1, H, W, generator=g)
  ws = [torch.randn(16, c, 7, 7, generator=g) * 0.1 for c in (3, 3, 1)]
  sh = torch.randn(3, 16, generator=g) * 0.2
  wst = torch.zeros(49, 7, 16)
  for w, c0 in zip(ws, (0, 3, 6)):
    wst[:, c0:c0 + w.shape[1]] = w.permute(2, 3, 1, 0).reshape(49, w.shape[1], 16)
  ref_all = sum(F.relu(F.conv2d(t, w, s, 1, 3)) for t, w, s in zip((img, pre, hm), ws, sh))
  ref_img = F.relu(F.conv2d(img, ws[0], sh[0], 1, 3))
  out = torch.empty(B, H, W, 16, device='cuda')
  wd, sd_, di, dp, dh = wst.cuda(), sh.cuda(), img.cuda(), pre.cuda(), hm.cuda()     # keep alive
  args = (L.ptr(wd), L.ptr(sd_), L.ptr(out), L.CT_F32, B, H, W, 16, L.stream_ptr())
  L.check(lib.ct_stem_forward(L.ptr(di), L.ptr(dp), L.ptr(dh), *args))
  assert (out.permute(0, 3, 1, 2).cpu() - ref_all).abs().max() < 2e-5
  L.check(lib.ct_stem_forward(L.ptr(di), L.ptr(None), L.ptr(None), *args))          # first frame of a plain detector
  assert (out.permute(0, 3, 1, 2).cpu() - ref_img).abs().max() < 2e-5


def test_conv_argument_validation():
  import ctypes as C
  lib = L.lib()
  d = L.ConvDesc()
  assert lib.ct_conv_forward(C.byref(d), None) == -1
  assert b'null pointer' in lib.ct_last_error()
