"""`Detector` with the reference's surface (src/lib/detector.py:23-458): `Detector(opt)`,
`run(image_or_path_or_tensor, meta={})`, `pre_process`, `process`, `post_process`, `merge_outputs`,
`reset_tracking`, attributes `.pause .tracker .pre_images .opt .model`; `run` returns the same dict
(`results` + the `tot/load/pre/net/dec/post/merge/track/display` wall-clock fields).

What differs underneath (`process`, detector.py:335-354): the network is a plan of libctb200 launches
(DLA34Engine) with the `_sigmoid_output` transforms fused into the head epilogues, decode is ONE fused
launch, and the 7-13 per-key blocking D2H copies become one copy of the packed record buffer.
Visualisation (`Debugger`, opt.debug >= 1) is outside the hot-path scope and is ignored.
"""
import copy
import math
import time

import numpy as np
import torch

from .dataset_info import get_dataset
from .decode import generic_decode
from .image import affine_transform, draw_umich_gaussian, gaussian_radius, get_affine_transform
from .model import create_model, load_model
from .post_process import generic_post_process
from .tracker import Tracker


class Detector(object):

  def __init__(self, opt):
    if opt.gpus[0] < 0 or not torch.cuda.is_available():
      raise RuntimeError('centertrack_b200.Detector needs a CUDA device (B200, sm_100a); there is no '
                         'CPU fallback (got --gpus %s)' % getattr(opt, 'gpus_str', opt.gpus))
    opt.device = torch.device('cuda')
    print('Creating model...')
    self.model = create_model(opt.arch, opt.heads, opt.head_conv, opt=opt)
    if opt.load_model != '':
      self.model = load_model(self.model, opt.load_model, opt)
    else:
      print('Warning: no --load_model given; running with randomly initialised weights')
    self.model = self.model.to(opt.device)
    self.model.eval()
    self.opt = opt
    self.trained_dataset = get_dataset(opt.dataset)
    self.mean = np.array(self.trained_dataset.mean, dtype=np.float32).reshape(1, 1, 3)
    self.std = np.array(self.trained_dataset.std, dtype=np.float32).reshape(1, 1, 3)
    self.pause = not opt.no_pause
    self.rest_focal_length = self.trained_dataset.rest_focal_length \
        if self.opt.test_focal_length < 0 else self.opt.test_focal_length
    self.flip_idx = self.trained_dataset.flip_idx
    self.cnt = 0
    self.pre_images = None
    self.pre_image_ori = None
    self.tracker = Tracker(opt)
    if opt.flip_test:
      raise NotImplementedError('--flip_test is scheduled after the main path (SURVEY 8f-3)')

  # ------------------------------------------------------------------------------------ run
  def run(self, image_or_path_or_tensor, meta={}):
    load_time, pre_time, net_time, dec_time, post_time = 0, 0, 0, 0, 0
    merge_time, track_time, tot_time, display_time = 0, 0, 0, 0
    start_time = time.time()
    pre_processed = False
    if isinstance(image_or_path_or_tensor, np.ndarray):
      image = image_or_path_or_tensor
    elif type(image_or_path_or_tensor) == type(''):
      import cv2
      image = cv2.imread(image_or_path_or_tensor)
    else:
      image = image_or_path_or_tensor['image'][0].numpy()
      pre_processed_images = image_or_path_or_tensor
      pre_processed = True
    loaded_time = time.time()
    load_time += (loaded_time - start_time)
    detections = []
    for scale in self.opt.test_scales:
      scale_start_time = time.time()
      if not pre_processed:
        images, meta = self.pre_process(image, scale, meta)
      else:
        images = pre_processed_images['images'][scale][0]
        meta = pre_processed_images['meta'][scale]
        meta = {k: v.numpy()[0] for k, v in meta.items()}
        if 'pre_dets' in pre_processed_images['meta']:
          meta['pre_dets'] = pre_processed_images['meta']['pre_dets']
        if 'cur_dets' in pre_processed_images['meta']:
          meta['cur_dets'] = pre_processed_images['meta']['cur_dets']
      images = images.to(self.opt.device, non_blocking=self.opt.non_block_test)
      pre_hms, pre_inds = None, None
      if self.opt.tracking:
        if self.pre_images is None:
          print('Initialize tracking!')
          self.pre_images = images
          self.tracker.init_track(meta['pre_dets'] if 'pre_dets' in meta else [])
        if self.opt.pre_hm:
          pre_hms, pre_inds = self._get_additional_inputs(
              self.tracker.tracks, meta, with_hm=not self.opt.zero_pre_hm)
      pre_process_time = time.time()
      pre_time += pre_process_time - scale_start_time
      output, dets, forward_time = self.process(images, self.pre_images, pre_hms, pre_inds,
                                                return_time=True)
      net_time += forward_time - pre_process_time
      decode_time = time.time()
      dec_time += decode_time - forward_time
      result = self.post_process(dets, meta, scale)
      post_process_time = time.time()
      post_time += post_process_time - decode_time
      detections.append(result)
    results = self.merge_outputs(detections)
    torch.cuda.synchronize()
    end_time = time.time()
    merge_time += end_time - post_process_time
    if self.opt.tracking:
      public_det = meta['cur_dets'] if self.opt.public_det else None
      results = self.tracker.step(results, public_det)
      self.pre_images = images
    tracking_time = time.time()
    track_time += tracking_time - end_time
    tot_time += tracking_time - start_time
    self.cnt += 1
    display_time += time.time() - end_time
    return {'results': results, 'tot': tot_time, 'load': load_time, 'pre': pre_time, 'net': net_time,
            'dec': dec_time, 'post': post_time, 'merge': merge_time, 'track': track_time,
            'display': display_time}

  # ------------------------------------------------------------------------------------ host pre
  def _transform_scale(self, image, scale=1):
    """detector.py:175-204."""
    import cv2
    height, width = image.shape[0:2]
    new_height, new_width = int(height * scale), int(width * scale)
    if self.opt.fix_short > 0:
      if height < width:
        inp_height = self.opt.fix_short
        inp_width = (int(width / height * self.opt.fix_short) + 63) // 64 * 64
      else:
        inp_height = (int(height / width * self.opt.fix_short) + 63) // 64 * 64
        inp_width = self.opt.fix_short
      c = np.array([width / 2, height / 2], dtype=np.float32)
      s = np.array([width, height], dtype=np.float32)
    elif self.opt.fix_res:
      inp_height, inp_width = self.opt.input_h, self.opt.input_w
      c = np.array([new_width / 2., new_height / 2.], dtype=np.float32)
      s = max(height, width) * 1.0
    else:
      inp_height = (new_height | self.opt.pad) + 1
      inp_width = (new_width | self.opt.pad) + 1
      c = np.array([new_width // 2, new_height // 2], dtype=np.float32)
      s = np.array([inp_width, inp_height], dtype=np.float32)
    resized_image = cv2.resize(image, (new_width, new_height))
    return resized_image, c, s, inp_width, inp_height, height, width

  def pre_process(self, image, scale, input_meta={}):
    """detector.py:207-239 (CPU only and fork-safe: test.py hands it to a DataLoader worker).
    Like the reference (hazard H5) `scale` is not forwarded to _transform_scale."""
    import cv2
    resized_image, c, s, inp_width, inp_height, height, width = self._transform_scale(image)
    trans_input = get_affine_transform(c, s, 0, [inp_width, inp_height])
    out_height = inp_height // self.opt.down_ratio
    out_width = inp_width // self.opt.down_ratio
    trans_output = get_affine_transform(c, s, 0, [out_width, out_height])
    inp_image = cv2.warpAffine(resized_image, trans_input, (inp_width, inp_height), flags=cv2.INTER_LINEAR)
    inp_image = ((inp_image / 255. - self.mean) / self.std).astype(np.float32)
    images = inp_image.transpose(2, 0, 1).reshape(1, 3, inp_height, inp_width)
    images = torch.from_numpy(images)
    meta = {'calib': np.array(input_meta['calib'], dtype=np.float32) if 'calib' in input_meta
            else self._get_default_calib(width, height)}
    meta.update({'c': c, 's': s, 'height': height, 'width': width, 'out_height': out_height,
                 'out_width': out_width, 'inp_height': inp_height, 'inp_width': inp_width,
                 'trans_input': trans_input, 'trans_output': trans_output})
    if 'pre_dets' in input_meta:
      meta['pre_dets'] = input_meta['pre_dets']
    if 'cur_dets' in input_meta:
      meta['cur_dets'] = input_meta['cur_dets']
    return images, meta

  def _trans_bbox(self, bbox, trans, width, height):
    bbox = np.array(copy.deepcopy(bbox), dtype=np.float32)
    bbox[:2] = affine_transform(bbox[:2], trans)
    bbox[2:] = affine_transform(bbox[2:], trans)
    bbox[[0, 2]] = np.clip(bbox[[0, 2]], 0, width - 1)
    bbox[[1, 3]] = np.clip(bbox[[1, 3]], 0, height - 1)
    return bbox

  def _get_additional_inputs(self, dets, meta, with_hm=True):
    """detector.py:254-290: render pre_hm at input resolution from the active tracks."""
    trans_input, trans_output = meta['trans_input'], meta['trans_output']
    inp_width, inp_height = meta['inp_width'], meta['inp_height']
    out_width, out_height = meta['out_width'], meta['out_height']
    input_hm = np.zeros((1, inp_height, inp_width), dtype=np.float32)
    output_inds = []
    for det in dets:
      if det['score'] < self.opt.pre_thresh or det['active'] == 0:
        continue
      bbox = self._trans_bbox(det['bbox'], trans_input, inp_width, inp_height)
      bbox_out = self._trans_bbox(det['bbox'], trans_output, out_width, out_height)
      h, w = bbox[3] - bbox[1], bbox[2] - bbox[0]
      if h > 0 and w > 0:
        radius = gaussian_radius((math.ceil(h), math.ceil(w)))
        radius = max(0, int(radius))
        ct = np.array([(bbox[0] + bbox[2]) / 2, (bbox[1] + bbox[3]) / 2], dtype=np.float32)
        ct_int = ct.astype(np.int32)
        if with_hm:
          draw_umich_gaussian(input_hm[0], ct_int, radius)
        ct_out = np.array([(bbox_out[0] + bbox_out[2]) / 2, (bbox_out[1] + bbox_out[3]) / 2],
                          dtype=np.int32)
        output_inds.append(ct_out[1] * out_width + ct_out[0])
    if with_hm:
      input_hm = torch.from_numpy(input_hm[np.newaxis]).to(self.opt.device)
    output_inds = np.array(output_inds, np.int64).reshape(1, -1)
    output_inds = torch.from_numpy(output_inds).to(self.opt.device)
    return input_hm, output_inds

  def _get_default_calib(self, width, height):
    return np.array([[self.rest_focal_length, 0, width / 2, 0],
                     [0, self.rest_focal_length, height / 2, 0],
                     [0, 0, 1, 0]])

  def _sigmoid_output(self, output):
    """detector.py:300-308 (kept for callers that run the nn.Module surface themselves; `process`
    fuses these transforms into the head epilogues)."""
    if 'hm' in output:
      output['hm'] = output['hm'].sigmoid_()
    if 'hm_hp' in output:
      output['hm_hp'] = output['hm_hp'].sigmoid_()
    if 'dep' in output:
      output['dep'] = 1. / (output['dep'].sigmoid() + 1e-6) - 1.
      output['dep'] *= self.opt.depth_scale
    return output

  # ------------------------------------------------------------------------------------ hot path
  def process(self, images, pre_images=None, pre_hms=None, pre_inds=None, return_time=False):
    """detector.py:335-354."""
    with torch.no_grad():
      torch.cuda.synchronize()
      B, _, H, W = images.shape
      eng = self.model.engine_for(B, H, W, images.device)
      if not eng.fused_act:
        eng.set_fused_activations(True)
      f = lambda t: None if t is None else t.float().contiguous()
      output = dict(eng.forward(f(images), f(pre_images), f(pre_hms) if isinstance(pre_hms, torch.Tensor)
                                else None))
      output.update({'pre_inds': pre_inds})
      torch.cuda.synchronize()
      forward_time = time.time()
      dets_dev = generic_decode(output, K=self.opt.K, opt=self.opt)
      rec = dets_dev.records.cpu().numpy()          # the single device->host copy (synchronises)
      dets = {}
      for k, v in dets_dev.items():
        if k == 'pre_cts':
          dets[k] = v.detach().cpu().numpy()
      dets.update(_numpy_views(rec, dets_dev))
    if return_time:
      return output, dets, forward_time
    return output, dets

  def post_process(self, dets, meta, scale=1):
    """detector.py:356-369."""
    dets = generic_post_process(self.opt, dets, [meta['c']], [meta['s']], meta['out_height'],
                                meta['out_width'], self.opt.num_classes, [meta['calib']],
                                meta['height'], meta['width'])
    self.this_calib = meta['calib']
    if scale != 1:
      for i in range(len(dets[0])):
        for k in ['bbox', 'hps']:
          if k in dets[0][i]:
            dets[0][i][k] = (np.array(dets[0][i][k], np.float32) / scale).tolist()
    return dets[0]

  def merge_outputs(self, detections):
    """detector.py:371-377."""
    assert len(self.opt.test_scales) == 1, 'multi_scale not supported!'
    return [d for d in detections[0] if d['score'] > self.opt.out_thresh]

  def reset_tracking(self):
    self.tracker.reset()
    self.pre_images = None
    self.pre_image_ori = None


def _numpy_views(rec, dets_dev):
  """Re-slice the host copy of the record buffer exactly like decode.views_from_records did on device."""
  from .decode import views_from_records
  t = torch.from_numpy(rec)
  out = views_from_records(t, dets_dev.layout)
  return {k: v.numpy() for k, v in out.items()}
