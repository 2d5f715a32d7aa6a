"""CPU oracle for the GaBOtorch geodesic-kernel / Riemannian-acquisition hot path.

THIS PACKAGE IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of
``bench.py`` may import it.  Nothing under ``gabotorch_b200/`` imports it, and the
product path raises when the CUDA library is missing instead of falling back here.

It restates, in plain numpy / torch-CPU fp64, the arithmetic of the reference
(`/root/reference`, GaBOtorch @ 884f64a), each function citing the file:line it follows:

* ``oracle.sphere``  – ``BoManifolds/Riemannian_utils/sphere_utils_torch.py:12-55``,
  ``kernel_utils/kernels_sphere.py:71-94,112-134`` and the numpy manifold formulas of
  ``Riemannian_utils/sphere_utils.py:14-123``.
* ``oracle.spd``     – ``Riemannian_utils/spd_utils_torch.py:13-226``,
  ``kernel_utils/kernels_spd.py:72-100,160-187,217-313`` and ``Riemannian_utils/spd_utils.py:57-306``.
* ``oracle.nested``  – ``nested_mappings/nested_spd_utils.py:13-118`` (projection and its approximate inverse) and
  ``kernel_utils/kernels_nested_spd.py``.
* ``oracle.nested_sphere`` – ``nested_mappings/nested_spheres_utils.py:13-213`` (both directions),
  ``Riemannian_utils/sphere_utils_torch.py:58-93``, ``kernel_utils/kernels_nested_sphere.py:129-152``.
* ``oracle.gp``      – the GP posterior / analytic Expected Improvement that the reference
  obtains from botorch/gpytorch (call sites ``examples/bo_sphere/benchmark_examples/gabo_sphere.py:131-165``), and the
  exact marginal log-likelihood ``fit_gpytorch_model`` optimises (``gabo_sphere.py:147,162``).
* ``oracle.rcg``     – the multi-start driver ``manifold_optimization/manifold_optimize.py:36-321``
  with the pymanopt 0.2.x ``ConjugateGradient`` + ``LineSearchAdaptive`` it calls.
* ``oracle.rtr``     – the reference's own trust-region solver ``manifold_optimization/robust_trust_regions.py:116-520``
  with the finite-difference Hessian ``manifold_optimization/approximate_hessian.py:11-62``.
* ``oracle.alm``     – the reference's own augmented Lagrangian solver
  ``manifold_optimization/augmented_Lagrange_method.py:66-328`` around its ``TrustRegions``.
* ``oracle.ctr``     – the reference's own constrained trust-region solvers (plain and strict)
  ``manifold_optimization/constrained_trust_regions.py:75-1415`` with the eigenvalue constraints of
  ``Riemannian_utils/spd_constraints_utils_torch.py:17-50`` (the configuration of ``gabo_spd.py``).

Parity pinning
--------------
PINNED (against the reference's own code imported from ``/root/reference`` with a
``torch.symeig`` shim, see ``tests/golden/make_golden.py`` and the committed fixtures):
sphere distance / kernel, Mandel pack/unpack, SPD affine-invariant distance / kernel,
Frobenius and log-Euclidean distance, nested SPD projection and reconstruction, ``sqrtm_torch``, nested-sphere
projection chain in both directions, and the trust-region solvers (the reference's ``TrustRegions`` and
``ConstrainedTrustRegions`` / ``StrictConstrainedTrustRegions`` / ``AugmentedLagrangeMethod`` classes themselves are run
by ``make_golden.py``; only its third-party base class ``pymanopt.solvers.solver.Solver`` -- the stopping rule -- is a
stand-in restated from pymanopt 0.2.x).

PARITY UNPINNED: everything whose arithmetic lives in pymanopt / botorch / gpytorch
(third-party, unpinned versions, absent from ``/root/reference`` and from this image):
manifold exp/log/retr/transp, conjugate gradient, line search, GP posterior, EI, marginal likelihood + priors.
Those are restated from the published algorithms (pymanopt 0.2.x, botorch 0.1-0.3,
gpytorch 1.x) and pinned only by mathematical identities and by the in-repo numpy
formulas (``sphere_utils.py``, ``spd_utils.py``) where they exist.
"""
