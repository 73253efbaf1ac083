"""Drop-in proof: runs the REFERENCE's own test-suite against this repo's `last_torch` alias.

    python tools/run_reference_tests.py --prepare     # in the build container: copies
        /root/reference/tests/*.py into _reference_tests/ (git-ignored, never committed;
        the directory travels to the GPU box with the gpurun snapshot) and writes the harness
        conftest there
    python tools/run_reference_tests.py [--cpu] [pytest args]   # runs them (GPU box: default
        device cuda; --cpu: host-side tests only, for iterating without a GPU)

The harness conftest does four things and edits no test: (0) seeds the global generators per test
(the tests draw unseeded random inputs); (1) puts the repo root first on
sys.path so that `import last_torch` is the alias package; (2) makes tensors the tests create land
on the GPU (torch.set_default_device + the legacy default tensor type for `torch.Tensor([...])`)
and lets numpy.testing read CUDA tensors (Tensor.__array__ via .cpu()); (3) marks the tests in
XFAIL below -- each one cites the reference defect or the deliberate
deviation that makes it fail -- as expected failures.
"""
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DEST = os.path.join(ROOT, '_reference_tests')
REFERENCE = os.environ.get('LAST_TORCH_REFERENCE', '/root/reference')

CONFTEST = r'''"""Harness for running the reference's tests against the `last_torch` alias
(written by tools/run_reference_tests.py --prepare; not part of the reference)."""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tools'))

import torch  # noqa: E402

USE_CUDA = os.environ.get('LT_REFTEST_DEVICE', 'cuda') == 'cuda' and torch.cuda.is_available()
if USE_CUDA:
  import warnings
  with warnings.catch_warnings():
    warnings.simplefilter('ignore')
    torch.set_default_tensor_type(torch.cuda.FloatTensor)    # torch.Tensor([...]) -> cuda
  torch.set_default_device('cuda')
  _orig_array = torch.Tensor.__array__

  def _array(self, dtype=None):
    return _orig_array(self.detach().cpu(), dtype) if dtype is not None else \
        _orig_array(self.detach().cpu())
  torch.Tensor.__array__ = _array
  _orig_numpy = torch.Tensor.numpy

  def _numpy(self, *args, **kwargs):          # the tests call .detach().numpy() on results
    return _orig_numpy(self.cpu() if self.is_cuda else self, *args, **kwargs)
  torch.Tensor.numpy = _numpy

import last_torch  # noqa: E402  (the alias)
assert 'last_torch_b200' in last_torch.RecognitionLattice.__module__

# The tests pin the labels of the reference AS SHIPPED (SURVEY D4 / D5): run shortest_path in its
# reference_compat mode (the default reports the true labels).
_orig_init = last_torch.RecognitionLattice.__init__


def _init(self, *args, **kwargs):
  _orig_init(self, *args, **kwargs)
  self.reference_compat = True
last_torch.RecognitionLattice.__init__ = _init

if USE_CUDA:
  # JointWeightFn's signature defaults to device='cpu' (weight_fns.py:187-192); the tests build
  # it without a device and feed it tensors of the default device
  _jw_init = last_torch.weight_fns.JointWeightFn.__init__

  def _jw(self, vocab_size, hidden_size, device='cuda', *args, **kwargs):
    _jw_init(self, vocab_size, hidden_size, device, *args, **kwargs)
  last_torch.weight_fns.JointWeightFn.__init__ = _jw

from run_reference_tests import XFAIL  # noqa: E402


@pytest.fixture(autouse=True)
def _reproducible_inputs(request):
  """The tests draw their inputs from the global generators without seeding them and compare
  fp32 results at numpy's default rtol = 1e-7: one ulp of summation-order difference (a device
  reduction instead of the CPU loop the expected value was written for) then decides pass or
  fail at random.  Seed per test, so that a run is reproducible."""
  import zlib
  seed = zlib.crc32(request.node.nodeid.encode()) & 0x7fffffff
  torch.manual_seed(seed)
  import numpy as np
  np.random.seed(seed)
  yield


def pytest_collection_modifyitems(config, items):
  for item in items:
    key = item.nodeid.split('::', 1)[-1] if '::' in item.nodeid else item.nodeid
    fname = os.path.basename(item.fspath)
    for (f, test), why in XFAIL.items():
      if f == fname and key.endswith(test):
        item.add_marker(pytest.mark.xfail(reason=why, strict=False))
'''

# (file, test id suffix) -> why it is expected to fail against this implementation.
XFAIL = {
    ('lattices_test.py', 'RecognitionLatticeCorrectnessTest::test_forward_backward'):
        'SURVEY D3: the test differentiates _forward / _forward_backward with torch.func.vjp '
        '(the reference\'s own _forward_backward has no working backward and the test compares '
        'torch.gradient of the OUTPUT vector at rtol=0.5, pinning nothing); the CUDA path exposes '
        'gradients through torch.autograd.Function nodes, which torch.func transforms reject, and '
        'the test ends in .numpy() on device tensors',
    ('weight_fns_test.py', 'SharedEmbCacher::test_call'):
        'SURVEY D7: the reference returns the nn.Embedding MODULE from the cacher and the test '
        'calls it; this implementation returns the [C, E] table that JointWeightFn needs '
        '(DESIGN.md section 1, deliberate deviation)',
}


def prepare():
  src = os.path.join(REFERENCE, 'tests')
  if not os.path.isdir(src):
    raise SystemExit(f'{src} not found (the reference only exists in the build container)')
  os.makedirs(DEST, exist_ok=True)
  for f in sorted(os.listdir(src)):
    if f.endswith('_test.py'):
      shutil.copy(os.path.join(src, f), os.path.join(DEST, f))
  with open(os.path.join(DEST, 'conftest.py'), 'w') as f:
    f.write(CONFTEST)
  print('prepared', DEST, sorted(os.listdir(DEST)))


def main():
  args = sys.argv[1:]
  if '--prepare' in args:
    prepare()
    return 0
  env = dict(os.environ)
  if '--cpu' in args:
    args.remove('--cpu')
    env['LT_REFTEST_DEVICE'] = 'cpu'
  if not os.path.isdir(DEST):
    raise SystemExit('run with --prepare first (in the build container)')
  with open(os.path.join(DEST, 'conftest.py'), 'w') as f:      # the harness is always current
    f.write(CONFTEST)
  cmd = [sys.executable, '-m', 'pytest', DEST, '-q', '-p', 'no:cacheprovider',
         '-o', 'python_files=*_test.py', '--rootdir', DEST] + args
  return subprocess.call(cmd, env=env, cwd=DEST)


if __name__ == '__main__':
  sys.exit(main())
