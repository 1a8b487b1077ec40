"""StreamRunner: the batched, graph-captured form of `Detector.process` for B independent video streams
on one GPU (streams shard across GPUs, replicas of the weights, no data-path collective).

One step = for every stream its current frame: network (pre_img = that stream's previous frame, kept on
the device exactly like detector.py:148) + fused sigmoid + fused decode -> one packed record buffer
[B,K,F].  THREE input slots rotate: step t reads slot t%3 (images) and slot (t-1)%3 (pre_images: the previous
step's images, never copied) while the copy stream uploads step t+1's frames into slot (t+1)%3 -- with two slots
the upload would have to wait for the step that still reads its target as pre_images.  One CUDA graph per slot
replays the whole step as a single launch.

The end-to-end form (`step_host`) takes HOST frames: pinned staging, H2D on a copy stream overlapped
with the previous step's compute, graph replay, D2H of the records.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib as L
from .decode import generic_decode


NS = 3          # input slots


class StreamRunner(object):

  def __init__(self, model, B, H, W, K=100, precision='bf16', device='cuda', use_graph=True):
    self.B, self.H, self.W, self.K = B, H, W, K
    self.device = torch.device(device)
    self.model = model
    self.eng = model.engine_for(B, H, W, self.device, precision)
    self.eng.set_fused_activations(True)
    f32 = torch.float32
    self.img = [torch.zeros((B, 3, H, W), dtype=f32, device=self.device) for _ in range(NS)]
    self.hm = [torch.zeros((B, 1, H, W), dtype=f32, device=self.device) for _ in range(NS)]
    self.h_img = [torch.zeros((B, 3, H, W), dtype=f32).pin_memory() for _ in range(NS)]
    self.h_hm = [torch.zeros((B, 1, H, W), dtype=f32).pin_memory() for _ in range(NS)]
    self.rec = None
    self.layout = None
    self.use_graph = use_graph
    self.graphs = [None] * NS
    self.compute = torch.cuda.Stream(device=self.device)
    self.copy = torch.cuda.Stream(device=self.device)
    self.ev_in = [torch.cuda.Event() for _ in range(NS)]     # slot uploaded
    self.ev_done = [torch.cuda.Event() for _ in range(NS)]   # slot no longer read (neither as images nor pre_images)
    self.t = 0
    self._eager(0)                                           # sizes the record buffer
    torch.cuda.synchronize(self.device)
    self.h_rec = [torch.zeros_like(self.rec, device='cpu').pin_memory() for _ in range(2)]
    self.launches_per_step = self.eng.n_launches + 1

  # one step, eager launches on the current stream
  def _eager(self, slot):
    out = dict(self.eng.forward(self.img[slot], self.img[(slot - 1) % NS], self.hm[slot]))
    res = generic_decode(out, K=self.K, records_out=self.rec)
    if self.rec is None:
      self.rec, self.layout = res.records, res.layout
    return res

  def _graph(self, slot):
    if self.graphs[slot] is None:
      s = torch.cuda.Stream(device=self.device)
      s.wait_stream(torch.cuda.current_stream())
      with torch.cuda.stream(s):
        self._eager(slot)
      torch.cuda.current_stream().wait_stream(s)
      g = torch.cuda.CUDAGraph()
      with torch.cuda.graph(g):
        self._eager(slot)
      self.graphs[slot] = g
    return self.graphs[slot]

  def warm(self):
    for s in range(NS):
      if self.use_graph:
        self._graph(s)
      else:
        self._eager(s)
    torch.cuda.synchronize(self.device)

  def load_device_inputs(self, images, pre_hms, slot):
    self.img[slot].copy_(images)
    self.hm[slot].copy_(pre_hms)

  def step_device(self):
    """Inputs already resident in the slot buffers; runs on the current stream."""
    slot = self.t % NS
    if self.use_graph:
      self._graph(slot).replay()
    else:
      self._eager(slot)
    self.t += 1
    return self.rec

  def step_host(self, images, pre_hms):
    """images [B,3,H,W], pre_hms [B,1,H,W]: float32 HOST tensors (what Detector.pre_process /
    _get_additional_inputs produce).  Returns the records of the PREVIOUS call (None the first time) --
    a one-step software pipeline: this step's H2D overlaps the previous step's compute."""
    slot = self.t % NS
    src_img, src_hm = images, pre_hms
    if not (images.is_pinned() and pre_hms.is_pinned()):   # pageable input: stage through pinned memory
      self.h_img[slot].copy_(images)
      self.h_hm[slot].copy_(pre_hms)
      src_img, src_hm = self.h_img[slot], self.h_hm[slot]
    with torch.cuda.stream(self.copy):
      self.copy.wait_event(self.ev_done[slot])       # slot's old contents were last read as pre_images of step t-2
      self.img[slot].copy_(src_img, non_blocking=True)
      self.hm[slot].copy_(src_hm, non_blocking=True)
      self.ev_in[slot].record(self.copy)
    prev = self.fetch() if self.t > 0 else None
    with torch.cuda.stream(self.compute):
      self.compute.wait_event(self.ev_in[slot])
      if self.use_graph:
        self._graph(slot).replay()
      else:
        self._eager(slot)
      self.h_rec[self.t & 1].copy_(self.rec, non_blocking=True)
      # this step's pre_images slot may be overwritten once this step is done
      self.ev_done[(slot - 1) % NS].record(self.compute)
    self.t += 1
    return prev

  def fetch(self, copy=True):
    """Blocks until the last submitted step finished; returns its records as numpy [B,K,F].  A copy by default: the
    two pinned buffers are reused by the asynchronous D2H of later steps (copy=False hands out the pinned view, valid
    until the step after next is submitted)."""
    self.compute.synchronize()
    rec = self.h_rec[(self.t - 1) & 1].numpy()
    return rec.copy() if copy else rec

  @property
  def h2d_bytes_per_step(self):
    return self.B * 4 * self.H * self.W * 4

  @property
  def d2h_bytes_per_step(self):
    return self.rec.numel() * 4

  def views(self, rec_np):
    from .decode import views_from_records
    return {k: v.numpy() for k, v in views_from_records(torch.from_numpy(rec_np), self.layout).items()}
