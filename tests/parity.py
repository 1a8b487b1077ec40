"""Parity metrics of one full-size frame against a reference golden (tests/golden/e2e_*.npz, written by
oracle/gen_golden.py from the UNMODIFIED reference).  Test infrastructure: used by the -m gpu parity tests, by
__graft_entry__.smoke() and by bench.py's `parity` key -- never by the product path.

All errors are ABSOLUTE errors divided by max(1, |ref|max) of the tensor they belong to: the scale north_star's
"fp32 heatmaps/offsets within 1e-3" is read in (tests/test_gpu_net.py, fp32 engine)."""
import numpy as np


def _np(t):
  return t.detach().float().cpu().numpy() if hasattr(t, 'detach') else np.asarray(t, dtype=np.float32)


def head_metrics(output, golden, frame=0):
  """output: {head: [B,c,h,w]} POST-activation maps (hm / hm_hp sigmoided, dep transformed).
  -> {head: {'max': max|d|/scale, 'rms': rms(d)/rms(ref)}} at the golden's 512 sampled positions."""
  pos = golden['pos']
  res = {}
  for k in golden.files:
    if not k.startswith('sample.'):
      continue
    h = k[len('sample.'):]
    if h not in output or output[h] is None:
      continue
    ref = golden[k]
    got = _np(output[h][frame]).reshape(ref.shape[0], -1)[:, pos]
    d = got - ref
    scale = max(1.0, float(np.abs(ref).max()))
    res[h] = {'max': float(np.abs(d).max() / scale),
              'rms': float(np.sqrt((d ** 2).mean()) / max(np.sqrt((ref ** 2).mean()), 1e-12))}
  return res


def peak_metrics(output, dets, golden, frame=0):
  """What bf16 costs in DETECTIONS.  At the reference's top-K peaks: |d score| (post-sigmoid), |d bbox|, |d tracking|
  computed from the device maps at the reference's own indices (so a rank swap between near-equal scores does not
  masquerade as a regression error); plus the overlap of the two top-K sets of (class, cell)."""
  hm = _np(output['hm'][frame])
  C, h, w = hm.shape
  r_cls = golden['det.clses'][0].astype(np.int64)
  r_x = golden['det.xs'][0].astype(np.int64)
  r_y = golden['det.ys'][0].astype(np.int64)
  r_ind = r_y * w + r_x
  res = {'score_max': float(np.abs(hm[r_cls, r_y, r_x] - golden['det.scores'][0]).max())}
  if 'tracking' in output and 'det.tracking' in golden.files:
    tr = _np(output['tracking'][frame]).reshape(2, -1)[:, r_ind].T
    res['tracking_max'] = float(np.abs(tr - golden['det.tracking'][0]).max())
  if 'wh' in output and 'reg' in output and 'det.bboxes' in golden.files:
    reg = _np(output['reg'][frame]).reshape(2, -1)[:, r_ind].T
    wh = np.maximum(_np(output['wh'][frame]).reshape(2, -1)[:, r_ind].T, 0)
    cx, cy = r_x + reg[:, 0], r_y + reg[:, 1]
    bb = np.stack([cx - wh[:, 0] / 2, cy - wh[:, 1] / 2, cx + wh[:, 0] / 2, cy + wh[:, 1] / 2], 1)
    res['bbox_max'] = float(np.abs(bb - golden['det.bboxes'][0]).max())
  if dets is not None:
    g_cls = np.asarray(dets['clses'])[frame].astype(np.int64)
    g_ind = (np.asarray(dets['ys'])[frame] * w + np.asarray(dets['xs'])[frame]).astype(np.int64)
    ref_set = set((r_cls * h * w + r_ind).tolist())
    got_set = set((g_cls * h * w + g_ind).tolist())
    res['topk_overlap'] = len(ref_set & got_set) / float(len(ref_set))
  return res


def stage_metrics(stage_fn, golden, frame=0):
  """stage_fn(name) -> NCHW fp32 tensor of a named intermediate.  -> {name: {'max', 'rms'}}."""
  res = {}
  for k in golden.files:
    if not k.startswith('stage.'):
      continue
    name = k[len('stage.'):]
    ref = golden[k]
    t = stage_fn('feat' if name == 'ida_up.node_2' else name)
    got = _np(t[frame]).reshape(ref.shape[0], -1)[:, golden['stagepos.' + name]]
    d = got - ref
    scale = max(1.0, float(np.abs(ref).max()))
    res[name] = {'max': float(np.abs(d).max() / scale),
                 'rms': float(np.sqrt((d ** 2).mean()) / max(np.sqrt((ref ** 2).mean()), 1e-12))}
  return res


def summarize(heads, peaks, stages=None):
  """One flat dict for the bench JSON / assertion messages."""
  out = {'head_max': {k: round(v['max'], 6) for k, v in heads.items()},
         'head_rms': {k: round(v['rms'], 6) for k, v in heads.items()}}
  out.update({k: round(v, 6) for k, v in peaks.items()})
  if stages:
    out['stage_max'] = {k: round(v['max'], 6) for k, v in stages.items()}
  return out
