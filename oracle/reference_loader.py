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


def load_trust_regions():
    """The reference's own ``TrustRegions`` class and ``get_hessianfd`` (manifold_optimization/robust_trust_regions.py,
    approximate_hessian.py).  The module imports the third-party base class ``pymanopt.solvers.solver.Solver``
    (absent here); a stand-in with pymanopt 0.2.x's published constructor defaults and stopping rule
    (``_check_stopping_criterion``: time, iter >= maxiter, gradnorm < mingradnorm, stepsize < minstepsize,
    costevals >= maxcostevals, in that order) is registered under that name first."""
    if not available():
        raise RuntimeError('reference tree not present at %s' % REFERENCE_ROOT)
    import importlib
    import time
    import types
    if 'pymanopt.solvers.solver' not in sys.modules:
        class Solver(object):
            def __init__(self, maxtime=1000, maxiter=1000, mingradnorm=1e-6, minstepsize=1e-10, maxcostevals=5000,
                         logverbosity=0):
                self._maxtime, self._maxiter, self._mingradnorm = maxtime, maxiter, mingradnorm
                self._minstepsize, self._maxcostevals, self._logverbosity = minstepsize, maxcostevals, logverbosity
                self._optlog = None

            def _check_stopping_criterion(self, time0, iter=-1, gradnorm=float('inf'), stepsize=float('inf'),
                                          costevals=-1):
                self._last_iter = iter               # kept so that the golden generator can record the iteration count
                if time.time() >= time0 + self._maxtime:
                    return 'max time'
                if iter >= self._maxiter:
                    return 'max iter'
                if gradnorm < self._mingradnorm:
                    return 'min grad norm'
                if stepsize < self._minstepsize:
                    return 'min stepsize'
                if costevals >= self._maxcostevals:
                    return 'max cost evals'
                return None

            def _start_optlog(self, *args, **kwargs):
                pass

            def _stop_optlog(self, *args, **kwargs):
                pass

        pkg = types.ModuleType('pymanopt')
        solvers = types.ModuleType('pymanopt.solvers')
        solver = types.ModuleType('pymanopt.solvers.solver')
        solver.Solver = Solver
        pkg.solvers = solvers
        solvers.solver = solver
        sys.modules.setdefault('pymanopt', pkg)
        sys.modules.setdefault('pymanopt.solvers', solvers)
        sys.modules['pymanopt.solvers.solver'] = solver
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    rtr = importlib.import_module('BoManifolds.manifold_optimization.robust_trust_regions')
    fd = importlib.import_module('BoManifolds.manifold_optimization.approximate_hessian')
    return rtr.TrustRegions, fd.get_hessianfd


def load_constrained_trust_regions():
    """The reference's own ``ConstrainedTrustRegions`` (manifold_optimization/constrained_trust_regions.py), its
    ``Problem`` with the PyTorch autodiff backend (pymanopt_addons, in-repo) and the eigenvalue constraints
    (Riemannian_utils/spd_constraints_utils_torch.py).  Needs the ``Solver`` stand-in of ``load_trust_regions`` and the
    ``symeig`` shim of ``load``."""
    load()
    load_trust_regions()
    import importlib
    ctr = importlib.import_module('BoManifolds.manifold_optimization.constrained_trust_regions')
    cons = importlib.import_module('BoManifolds.Riemannian_utils.spd_constraints_utils_torch')
    fd = importlib.import_module('BoManifolds.manifold_optimization.approximate_hessian')
    ctr.ConstrainedTrustRegions.Strict = ctr.StrictConstrainedTrustRegions      # handed out together
    return ctr.ConstrainedTrustRegions, fd.get_hessianfd, cons


def load_alm():
    """The reference's own ``AugmentedLagrangeMethod`` (manifold_optimization/augmented_Lagrange_method.py).  It does
    ``import pymanopt`` and tests ``isinstance(inner_solver, pymanopt.solvers.NelderMead)``: the stand-in package of
    ``load_trust_regions`` gets an empty ``NelderMead`` class for that test."""
    load()
    load_trust_regions()
    import importlib
    solvers = sys.modules['pymanopt.solvers']
    if not hasattr(solvers, 'NelderMead'):
        solvers.NelderMead = type('NelderMead', (), {})
    alm = importlib.import_module('BoManifolds.manifold_optimization.augmented_Lagrange_method')
    return alm.AugmentedLagrangeMethod


def load_nested_spd_optimization():
    """The reference's reconstruction costs (nested_mappings/nested_spd_optimization.py:22-92).  The module imports
    ``gpytorch`` and ``pymanopt.manifolds`` at the top (used only inside ``optimize_reconstruction_parameters_nested_spd``,
    which needs the absent pymanopt ``Product`` / ``Grassmann`` manifolds and is NOT run): empty stand-in modules are
    registered for the import; the two cost functions themselves run unmodified reference code."""
    load_alm()
    import importlib
    import types
    if 'gpytorch' not in sys.modules:
        sys.modules['gpytorch'] = types.ModuleType('gpytorch')
    pkg = sys.modules['pymanopt']
    if 'pymanopt.manifolds' not in sys.modules:
        man = types.ModuleType('pymanopt.manifolds')
        pkg.manifolds = man
        sys.modules['pymanopt.manifolds'] = man
    return importlib.import_module('BoManifolds.nested_mappings.nested_spd_optimization')
