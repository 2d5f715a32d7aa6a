"""Import the importable slices of the actual reference from ``/root/reference``.  Test infrastructure only.

Only usable in the build container (``/root/reference`` does not exist on the GPU box); it is used by
``tests/golden/make_golden.py`` to generate the committed fixtures and by the ``not gpu`` tests that
re-check the oracle against live reference code when the tree is present.

torch 2.11 removed ``torch.symeig``, which the reference calls (``spd_utils_torch.py:25,45,110``); the shim
below maps it onto ``torch.linalg.eigh(UPLO='U')`` (``symeig``'s default was ``upper=True``).
"""
import collections
import os
import sys
import warnings

REFERENCE_ROOT = '/root/reference'


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, 'BoManifolds'))


def load():
    """Returns a namespace with the reference modules that import only torch/numpy/scipy."""
    if not available():
        raise RuntimeError('reference tree not present at %s' % REFERENCE_ROOT)
    import torch
    if not hasattr(torch, '_gabo_symeig_shim'):
        _Sym = collections.namedtuple('symeig', ['eigenvalues', 'eigenvectors'])

        def symeig(x, eigenvectors=False, upper=True):
            lam, vec = torch.linalg.eigh(x, UPLO='U' if upper else 'L')
            return _Sym(lam, vec)
        torch.symeig = symeig
        torch._gabo_symeig_shim = True
    warnings.filterwarnings('ignore', message='.*torch.cholesky is deprecated.*')
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import importlib
    ns = collections.OrderedDict()
    ns['sphere_utils_torch'] = importlib.import_module('BoManifolds.Riemannian_utils.sphere_utils_torch')
    ns['spd_utils_torch'] = importlib.import_module('BoManifolds.Riemannian_utils.spd_utils_torch')
    ns['sphere_utils'] = importlib.import_module('BoManifolds.Riemannian_utils.sphere_utils')
    ns['spd_utils'] = importlib.import_module('BoManifolds.Riemannian_utils.spd_utils')
    ns['nested_spd_utils'] = importlib.import_module('BoManifolds.nested_mappings.nested_spd_utils')
    ns['nested_spheres_utils'] = importlib.import_module('BoManifolds.nested_mappings.nested_spheres_utils')
    return type('Reference', (), dict(ns))
