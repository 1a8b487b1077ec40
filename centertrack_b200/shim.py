"""install(): make the reference's import names resolve to this package, so the reference's own
`src/demo.py` / `src/test.py` (which do `from detector import Detector`, `from opts import opts`) and any
code importing `model.model`, `model.decode`, `model.networks.DCNv2.dcn_v2`, `utils.tracker`,
`utils.post_process`, `utils.image` run unchanged on the B200 path.  See INTEGRATION.md.

    import centertrack_b200.shim as shim; shim.install()        # before importing the reference scripts
    # or, DCN only (keep the reference's PyTorch graph, swap its absent CUDA extension):
    shim.install_dcn_only()
"""
import sys
import types


def _pkg(name):
  m = sys.modules.get(name)
  if m is None:
    m = types.ModuleType(name)
    m.__path__ = []
    sys.modules[name] = m
  return m


def install_dcn_only():
  """Register `model.networks.DCNv2.dcn_v2.DCN` (dla.py:18-22 imports it inside a try/except) without
  touching anything else: the reference's own DLASeg then runs with our DCN kernels."""
  from . import dcn
  _pkg('model.networks.DCNv2')
  mod = types.ModuleType('model.networks.DCNv2.dcn_v2')
  mod.DCN = dcn.DCN
  sys.modules['model.networks.DCNv2.dcn_v2'] = mod
  return mod


def install():
  from . import dataset_info, decode, detector, image, model, opts, post_process, tracker
  alias = {
      'detector': detector, 'opts': opts,
      'model.model': model, 'model.decode': decode,
      'utils.tracker': tracker, 'utils.post_process': post_process, 'utils.image': image,
      'dataset.dataset_factory': dataset_info,
  }
  for pkg in ('model', 'model.networks', 'utils', 'dataset'):
    _pkg(pkg)
  for name, mod in alias.items():
    sys.modules[name] = mod
  install_dcn_only()
  return alias
