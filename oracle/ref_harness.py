"""TEST INFRASTRUCTURE ONLY -- never imported by the product path.

Loads the *unmodified* reference (xingyizhou/CenterTrack, /root/reference/src/lib) in
this container on CPU so that golden vectors can be generated from the reference's own
Python and the oracle restatement (oracle/ct_oracle.py) can be pinned against it.

/root/reference does not exist on the GPU box: nothing that runs there imports this
module (tests that need it are skipped when REF_ROOT is absent).

What is shimmed (all off the arithmetic path):
  * stub modules for absent packages: progress.bar, pycocotools.{coco,cocoeval},
    pyquaternion, matplotlib(.pyplot), mpl_toolkits.mplot3d,
    sklearn.utils.linear_assignment_ (removed API, tracker.py:2),
    torchvision.models.utils (removed API, backbones/mobilenet.py:14)
  * model.networks.DCNv2.dcn_v2.DCN: the DCNv2 submodule is EMPTY in the reference
    checkout (un-vendored, CharlesShang/DCNv2 @ master, no SHA). Stand-in =
    torchvision.ops.deform_conv2d with identical parameter names
    (weight, bias, conv_offset_mask.{weight,bias}).  => DCN parity is "unpinned" by
    the reference itself; it is anchored on torchvision + the independent restatement
    in ct_oracle.dcn_v2_forward (Appendix B semantics).
  * torch.cuda.synchronize -> no-op when no driver (detector.py:139,338,344,348).
"""
import os
import sys
import types
import math

import numpy as np
import torch
import torch.nn as nn

REF_ROOT = os.environ.get('CT_REF_ROOT', '/root/reference')
REF_LIB = os.path.join(REF_ROOT, 'src', 'lib')


def available():
  return os.path.isdir(REF_LIB)


class _RefDCN(nn.Module):
  """Stand-in for model/networks/DCNv2/dcn_v2.py::DCN (absent submodule).

  Follows upstream DCNv2 semantics: conv_offset_mask ->
  27 ch; o1,o2,mask = chunk(3); offset = cat(o1,o2); mask = sigmoid(mask).
  """

  def __init__(self, in_channels, out_channels, kernel_size=(3, 3), stride=1,
               padding=1, dilation=1, deformable_groups=1):
    super().__init__()
    kh, kw = kernel_size if isinstance(kernel_size, (tuple, list)) \
        else (kernel_size, kernel_size)
    self.stride, self.padding, self.dilation = stride, padding, dilation
    self.weight = nn.Parameter(torch.empty(out_channels, in_channels, kh, kw))
    self.bias = nn.Parameter(torch.zeros(out_channels))
    self.conv_offset_mask = nn.Conv2d(
        in_channels, deformable_groups * 3 * kh * kw, kernel_size=(kh, kw),
        stride=stride, padding=padding, bias=True)
    n = in_channels * kh * kw
    stdv = 1. / math.sqrt(n)
    self.weight.data.uniform_(-stdv, stdv)
    self.conv_offset_mask.weight.data.zero_()
    self.conv_offset_mask.bias.data.zero_()

  def forward(self, x):
    from torchvision.ops import deform_conv2d
    out = self.conv_offset_mask(x)
    o1, o2, mask = torch.chunk(out, 3, dim=1)
    offset = torch.cat((o1, o2), dim=1)
    mask = torch.sigmoid(mask)
    return deform_conv2d(x, offset, self.weight, self.bias, self.stride,
                         self.padding, self.dilation, mask)


def _stub(name, **attrs):
  m = types.ModuleType(name)
  for k, v in attrs.items():
    setattr(m, k, v)
  sys.modules[name] = m
  return m


_installed = False


def install_third_party_stubs():
  """Stub modules for the packages the reference imports but this image lacks (all off the arithmetic path)."""

  class _Bar(object):
    suffix = ''
    elapsed_td = eta_td = 0
    def __init__(self, *a, **k): pass
    def next(self): pass
    def finish(self): pass

  if 'progress' not in sys.modules:
    _stub('progress'); _stub('progress.bar', Bar=_Bar)
  try:
    import pycocotools.coco  # noqa
  except Exception:
    _stub('pycocotools'); _stub('pycocotools.coco', COCO=object)
    _stub('pycocotools.cocoeval', COCOeval=object)
  try:
    import pyquaternion  # noqa
  except Exception:
    _stub('pyquaternion', Quaternion=object)
  try:
    import matplotlib.pyplot  # noqa
  except Exception:
    _stub('matplotlib'); _stub('matplotlib.pyplot')
    _stub('mpl_toolkits'); _stub('mpl_toolkits.mplot3d', Axes3D=object)
  try:
    from sklearn.utils.linear_assignment_ import linear_assignment  # noqa
  except Exception:
    def linear_assignment(cost):
      from scipy.optimize import linear_sum_assignment
      r, c = linear_sum_assignment(cost)
      return np.stack([r, c], axis=1)
    _stub('sklearn.utils.linear_assignment_', linear_assignment=linear_assignment)
  try:
    import torchvision.models.utils  # noqa
  except Exception:
    _stub('torchvision.models.utils',
          load_state_dict_from_url=torch.hub.load_state_dict_from_url)


def install():
  """Make `import detector`, `import model.model`, ... resolve to the reference."""
  global _installed
  if _installed:
    return
  if not available():
    raise RuntimeError('reference checkout not present at %s' % REF_ROOT)
  install_third_party_stubs()
  pkg = _stub('model.networks.DCNv2')
  pkg.__path__ = []
  _stub('model.networks.DCNv2.dcn_v2', DCN=_RefDCN)
  if REF_LIB not in sys.path:
    sys.path.insert(0, REF_LIB)
  # `model` must resolve to the reference package; re-register the DCN stub below it.
  import model.networks  # noqa
  sys.modules['model.networks.DCNv2'] = pkg
  sys.modules['model.networks.DCNv2.dcn_v2'] = sys.modules['model.networks.DCNv2.dcn_v2']
  if not torch.cuda.is_available():
    torch.cuda.synchronize = lambda *a, **k: None
  _installed = True


def no_pretrained_download():
  """--arch generic builds its backbone with `dla34(opt=opt)` whose `pretrained=True` default fetches ImageNet weights
  from dl.yf.io (backbones/dla.py:362-371,305-316; SURVEY hazard H7) -- there is no network here and every weight
  is overwritten by the synthetic checkpoint anyway, so the download (and only the download) is skipped."""
  install()
  import model.networks.backbones.dla as bdla
  bdla.DLA.load_pretrained_model = lambda self, *a, **k: None


TASK_ARGS = {
    # name -> (task, extra argv)   (BASELINE.json configs 1..5)
    'coco_tracking': ('tracking', []),
    'mot': ('tracking', ['--num_classes', '1', '--input_h', '544', '--input_w', '960']),
    'nuscenes_ddd': ('tracking,ddd', []),
    'coco_pose': ('tracking,multi_pose', []),
}


def make_opt(cfg='coco_tracking', load_model='', extra=(), input_hw=None):
  """Build `opt` through the reference's own opts().init() (opts.py:390-403)."""
  install()
  from opts import opts
  task, argv = TASK_ARGS[cfg]
  argv = list(argv)
  if input_hw is not None:
    argv = [a for a in argv]
    # override resolution
    argv += ['--input_h', str(input_hw[0]), '--input_w', str(input_hw[1])]
  old = sys.argv
  sys.argv = ['demo.py', task, '--gpus', '-1', '--load_model', load_model,
              '--pre_hm', '--track_thresh', '0.01', '--new_thresh', '0.01'] + argv + list(extra)
  try:
    import io, contextlib
    with contextlib.redirect_stdout(io.StringIO()):
      opt = opts().init()
  finally:
    sys.argv = old
  opt.debug = 0
  return opt


def he_init_(model, seed=317, hm_scale=0.25):
  """Variance-preserving seeded init: reference default init collapses
  the feature to ~1e-5 (hazard H3) so every top-K would be a tie."""
  g = torch.Generator().manual_seed(seed)
  with torch.no_grad():
    for name, m in model.named_modules():
      if isinstance(m, nn.Conv2d):
        if name.endswith('conv_offset_mask'):
          m.weight.normal_(0, 0.01, generator=g)
          m.bias.zero_()
        else:
          fan_in = m.weight.shape[1] * m.weight.shape[2] * m.weight.shape[3]
          m.weight.normal_(0, math.sqrt(2.0 / fan_in), generator=g)
          if m.bias is not None:
            m.bias.zero_()
      elif isinstance(m, _RefDCN):
        fan_in = m.weight.shape[1] * 9
        m.weight.normal_(0, math.sqrt(2.0 / fan_in), generator=g)
        m.bias.normal_(0, 0.05, generator=g)
      elif isinstance(m, nn.BatchNorm2d):
        # non-trivial but benign BN statistics so BN folding is actually exercised
        m.weight.uniform_(0.8, 1.2, generator=g)
        m.bias.normal_(0, 0.05, generator=g)
        m.running_mean.normal_(0, 0.05, generator=g)
        m.running_var.uniform_(0.8, 1.2, generator=g)
    for head in model.heads:
      fc = getattr(model, head)
      if 'hm' in head:
        fc[-1].weight.mul_(hm_scale)
        fc[-1].bias.fill_(-4.6)
      else:
        fc[-1].bias.normal_(0, 0.1, generator=g)
  return model


def build_reference_model(cfg='coco_tracking', seed=317, ckpt_path=None, input_hw=None, extra=()):
  """create_model (model.py:24-29) + He init + save_model (model.py:92-101)."""
  install()
  import io, contextlib
  from model.model import create_model, save_model
  opt = make_opt(cfg, load_model='dummy.pth', input_hw=input_hw, extra=extra)
  with contextlib.redirect_stdout(io.StringIO()):
    model = create_model(opt.arch, opt.heads, opt.head_conv, opt=opt)
  he_init_(model, seed)
  model.eval()
  if ckpt_path is not None:
    save_model(ckpt_path, 0, model)
  return opt, model
