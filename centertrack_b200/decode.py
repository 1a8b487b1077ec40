"""`generic_decode(output, K=100, opt=None)` with the reference's contract (model/decode.py:83-182):
same keys, shapes and dtypes (all float32 tensors on `output`'s device), computed by ONE fused CUDA
launch (ct_decode: 3x3 max-equals NMS + per-class top-K + cross-class top-K + gather of every head +
box assembly + pose refinement) into one packed record buffer; the returned tensors are views of it,
so the caller's `dets[k].cpu()` loop costs a single D2H when it reads `result.records` instead.

Tie rule (the reference leaves ties to torch.topk): score descending, then flat index ascending.
"""
import ctypes as C

import torch

from . import _lib as L

_RAW_HEADS = ['tracking', 'dep', 'rot', 'dim', 'amodel_offset', 'nuscenes_att', 'velocity']
_ROLES = [('reg', L.CT_ROLE_REG), ('wh', L.CT_ROLE_WH), ('ltrb', L.CT_ROLE_LTRB)] + \
    [(h, L.CT_ROLE_RAW) for h in _RAW_HEADS] + \
    [('ltrb_amodal', L.CT_ROLE_LTRB_AMODAL), ('hps', L.CT_ROLE_HPS)]


class DecodeResult(dict):
  """dict of views + the packed buffer: `.records` [B,K,F] fp32, `.layout` {key: (offset, width)},
  `.inds` [B,K] int32 flat indices (bit-exact top-K evidence)."""
  records = None
  layout = None

  @property
  def inds(self):
    return self.records[:, :, L.CT_REC_IND].contiguous().view(torch.int32)


_ws_cache = {}


def _workspace(device, B, C_, J, K):
  """Scratch of one ct_decode launch (per-batch arrival counters + candidate lists).  Keyed by the CUDA stream as
  well: two runners / Detectors driving different streams must not share the counters."""
  import torch as _t
  key = (str(device), B, C_, J, K, int(_t.cuda.current_stream(device).cuda_stream))
  ws = _ws_cache.get(key)
  if ws is None:
    n = L.lib().ct_decode_workspace_bytes(B, C_, J, K)
    ws = torch.zeros(n, dtype=torch.uint8, device=device)   # zeroed once: arrival counters
    _ws_cache[key] = ws
  return ws


def _f32c(t, name):
  if not (t.is_cuda and t.dtype == torch.float32):
    raise RuntimeError('generic_decode: %r must be a float32 CUDA tensor (no CPU fallback), got %s %s'
                       % (name, t.dtype, t.device))
  return t if t.is_contiguous() else t.contiguous()


def generic_decode(output, K=100, opt=None, records_out=None, workspace=None):
  """`workspace`: optional caller-owned scratch tensor (ct_decode_workspace_bytes, zero-initialised once); a
  CUDA-graph capture must pass its own so that the captured pointer stays alive and private."""
  if 'hm' not in output:
    return {}
  if opt is not None and getattr(opt, 'zero_tracking', False) and 'tracking' in output:
    output['tracking'] *= 0
  heat = _f32c(output['hm'], 'hm')
  B, cat, H, W = heat.shape
  d = L.DecodeDesc()
  d.B, d.C, d.H, d.W, d.K = B, cat, H, W, K
  d.hm = heat.data_ptr()
  keep = [heat]
  layout = {}
  off = L.CT_REC_HEADS
  n = 0
  for name, role in _ROLES:
    if name not in output or output[name] is None:
      continue
    t = _f32c(output[name], name)
    assert t.shape[0] == B and t.shape[2:] == (H, W), (name, t.shape)
    if role in (L.CT_ROLE_REG, L.CT_ROLE_WH) and t.shape[1] != 2:
      raise NotImplementedError('%s head with %d channels (class-specific wh) is not supported'
                                % (name, t.shape[1]))
    keep.append(t)
    d.heads[n].map, d.heads[n].channels, d.heads[n].role, d.heads[n].rec_offset = \
        t.data_ptr(), t.shape[1], role, off
    layout[name] = (off, t.shape[1])
    off += t.shape[1]
    n += 1
  d.n_heads = n
  d.has_bbox = int(any(k in output for k in ('wh', 'ltrb', 'ltrb_amodal')))
  J = 0
  d.rec_hps = d.rec_kps_score = -1
  if 'hps' in output and 'hm_hp' in output:
    hp = _f32c(output['hm_hp'], 'hm_hp')
    keep.append(hp)
    J = hp.shape[1]
    assert output['hps'].shape[1] == 2 * J
    if not d.has_bbox:
      raise NotImplementedError('pose refinement without a box head is not supported')
    d.hm_hp, d.J = hp.data_ptr(), J
    if 'hp_offset' in output:
      ho = _f32c(output['hp_offset'], 'hp_offset')
      keep.append(ho)
      d.hp_offset = ho.data_ptr()
    d.rec_hps = off
    layout['hps_refined'] = (off, 2 * J)
    off += 2 * J
    d.rec_kps_score = off
    layout['kps_score'] = (off, 1)
    off += 1
  d.rec_floats = off
  if records_out is None:
    records_out = torch.empty((B, K, off), dtype=torch.float32, device=heat.device)
  else:
    assert records_out.shape == (B, K, off) and records_out.is_contiguous()
  d.records = records_out.data_ptr()
  ws = workspace if workspace is not None else _workspace(heat.device, B, cat, J, K)
  assert ws.numel() * ws.element_size() >= L.lib().ct_decode_workspace_bytes(B, cat, J, K)
  d.workspace = ws.data_ptr()
  L.check(L.lib().ct_decode(C.byref(d), L.stream_ptr()), 'ct_decode')
  return views_from_records(records_out, layout, output, W)


def views_from_records(rec, layout, output=None, W=None):
  """Slice the packed [B,K,F] buffer into the reference's ret dict (decode.py:97-181)."""
  ret = DecodeResult()
  ret.records, ret.layout = rec, layout
  ret['scores'] = rec[:, :, L.CT_REC_SCORE]
  ret['clses'] = rec[:, :, L.CT_REC_CLS]
  ret['xs'] = rec[:, :, L.CT_REC_XS]
  ret['ys'] = rec[:, :, L.CT_REC_YS]
  ret['cts'] = rec[:, :, L.CT_REC_XS:L.CT_REC_YS + 1]
  if any(k in layout for k in ('wh', 'ltrb', 'ltrb_amodal')):
    ret['bboxes'] = rec[:, :, L.CT_REC_BBOX:L.CT_REC_BBOX + 4]
  for h in _RAW_HEADS:
    if h in layout:
      o, w = layout[h]
      ret[h] = rec[:, :, o:o + w]
  if 'ltrb_amodal' in layout:
    ret['bboxes_amodal'] = ret['bboxes']
  if 'hps' in layout:
    if 'hps_refined' in layout:
      o, w = layout['hps_refined']
      ret['hps'] = rec[:, :, o:o + w]
      ret['kps_score'] = rec[:, :, layout['kps_score'][0]]
    else:                                   # decode.py:80-81: no hm_hp -> (kps, kps)
      o, w = layout['hps']
      ret['hps'] = rec[:, :, o:o + w]
      ret['kps_score'] = ret['hps']
  if output is not None and output.get('pre_inds', None) is not None:
    pre_inds = output['pre_inds']
    ret['pre_cts'] = torch.stack([(pre_inds % W).float(), torch.div(pre_inds, W, rounding_mode='floor').float()],
                                 dim=2)
  return ret
