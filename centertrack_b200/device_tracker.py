"""Per-stream tracker state kept on the GPU between frames: the reference's
`generic_post_process` affine (utils/post_process.py:21-91), `Tracker.step` greedy association
(utils/tracker.py:28-138) and the prior heat-map render of `Detector._get_additional_inputs`
(detector.py:254-290) as two launches on the decode records -- `ct_track_step` and `ct_render_tracks` -- so that a
stream never returns to the host between frames: records(t) -> tracks(t) -> pre_hm(t+1) are all device-resident and
CUDA-graph capturable (fixed launch shapes).

Greedy association only: `--hungarian` / `--public_det` streams use the host tracker (centertrack_b200.tracker,
`Detector.run`); this class refuses them rather than silently tracking differently.
Results are rows of CT_TRK_FLOATS fp32 (score, class, ct, tracking, bbox, tracking_id, age, active) in the
reference's output order (matched detections, new tracks, coasting tracks); `results()` turns a host copy into the
reference's list of dicts.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib as L
from .image import get_affine_transform


class DeviceTracker(object):

  def __init__(self, opt, B, K, rec_floats, layout, inp_h, inp_w, device, centers=None, scales=None, max_tracks=None):
    """centers/scales: per-stream (c, s) of the source rectangle (Detector._input_geometry); default = a source image
    of exactly the network input size (the synthetic benchmark streams)."""
    if getattr(opt, 'hungarian', False) or getattr(opt, 'public_det', False):
      raise NotImplementedError('the device tracker is greedy-only: --hungarian / --public_det run on the host '
                                '(centertrack_b200.tracker.Tracker via Detector.run)')
    self.opt, self.B, self.K, self.F = opt, B, K, rec_floats
    self.inp_h, self.inp_w = inp_h, inp_w
    self.device = torch.device(device)
    self.T = int(max_tracks or (K * (1 + max(0, int(opt.max_age)))))
    self.T = max(self.T, K)
    down = getattr(opt, 'down_ratio', 4)
    out_w, out_h = inp_w // down, inp_h // down
    t_out = np.zeros((B, 6), np.float32)
    t_in = np.zeros((B, 6), np.float64)
    for b in range(B):
      c = np.array([inp_w / 2., inp_h / 2.], np.float32) if centers is None else np.asarray(centers[b], np.float32)
      s = max(inp_h, inp_w) * 1.0 if scales is None else scales[b]
      t_out[b] = get_affine_transform(c, s, 0, (out_w, out_h), inv=1).astype(np.float32).reshape(6)
      t_in[b] = get_affine_transform(c, s, 0, [inp_w, inp_h]).reshape(6)
    self.trans_out_inv = torch.from_numpy(t_out).to(self.device)
    self.trans_input = torch.from_numpy(t_in).to(self.device)
    self.tracks = torch.zeros((B, self.T, L.CT_TRK_FLOATS), dtype=torch.float32, device=self.device)
    self.counts = torch.zeros((B, 2), dtype=torch.int32, device=self.device)
    self.boxes = torch.zeros((B, self.T, 5), dtype=torch.float32, device=self.device)
    self.boxes[:, :, 3] = -1.0                       # no tracks yet: nothing to splat (first frame: pre_hm = 0)
    d = L.TrackDesc()
    d.B, d.K, d.F = B, K, rec_floats
    d.rec_tracking = layout['tracking'][0] if 'tracking' in layout else -1
    d.max_tracks = self.T
    d.out_thresh, d.new_thresh, d.pre_thresh = float(opt.out_thresh), float(opt.new_thresh), float(opt.pre_thresh)
    d.max_age = int(opt.max_age)
    d.inp_h, d.inp_w = inp_h, inp_w
    d.trans_out_inv, d.trans_input = self.trans_out_inv.data_ptr(), self.trans_input.data_ptr()
    d.tracks, d.counts, d.boxes = self.tracks.data_ptr(), self.counts.data_ptr(), self.boxes.data_ptr()
    self.desc = d
    if L.lib().ct_track_smem_bytes(K, self.T) > 200 * 1024:
      raise ValueError('track table of %d rows does not fit in shared memory' % self.T)

  def reset(self):
    """Detector.reset_tracking / Tracker.reset for every stream."""
    self.tracks.zero_()
    self.counts.zero_()
    self.boxes.zero_()
    self.boxes[:, :, 3] = -1.0

  def step(self, records):
    """records [B,K,F] (ct_decode) -> updates tracks/counts/boxes in place, on the current stream."""
    assert records.is_cuda and records.dtype == torch.float32 and tuple(records.shape) == (self.B, self.K, self.F)
    self.desc.records = records.data_ptr()
    L.check(L.lib().ct_track_step(C.byref(self.desc), L.stream_ptr()), 'ct_track_step')

  def render(self, pre_hm):
    """pre_hm [B,1,H,W] fp32 <- splat of the current tracks (what the NEXT frame's network reads)."""
    assert pre_hm.is_cuda and pre_hm.dtype == torch.float32 and tuple(pre_hm.shape) == (self.B, 1, self.inp_h, self.inp_w)
    L.check(L.lib().ct_render_tracks(L.ptr(self.boxes), self.B * self.T, L.ptr(pre_hm), self.B, self.inp_h,
                                     self.inp_w, L.stream_ptr()), 'ct_render_tracks')

  @property
  def d2h_bytes(self):
    return self.tracks.numel() * 4 + self.counts.numel() * 4

  @staticmethod
  def results(tracks_np, counts_np):
    """Host copies -> [[{score, class, ct, tracking, bbox, tracking_id, age, active}, ...] per stream]."""
    out = []
    for b in range(tracks_np.shape[0]):
      rows = tracks_np[b, :int(counts_np[b, 0])]
      out.append([{'score': float(r[L.CT_TRK_SCORE]), 'class': int(r[L.CT_TRK_CLASS]),
                   'ct': r[L.CT_TRK_CT:L.CT_TRK_CT + 2].copy(), 'tracking': r[L.CT_TRK_TRACKING:L.CT_TRK_TRACKING + 2].copy(),
                   'bbox': r[L.CT_TRK_BBOX:L.CT_TRK_BBOX + 4].copy(), 'tracking_id': int(r[L.CT_TRK_ID]),
                   'age': int(r[L.CT_TRK_AGE]), 'active': int(r[L.CT_TRK_ACTIVE])} for r in rows])
    return out
