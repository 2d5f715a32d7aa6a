"""gabotorch_b200 -- B200-native (sm_100a) implementation of GaBOtorch's data-parallel hot path.

Geodesic-kernel Gram builds on S^d and SPD(d), the batched Riemannian acquisition optimiser and the nested SPD
projection, behind the reference's own class / function names.  Host code is Python/PyTorch (device memory, streams,
torch.distributed); the arithmetic runs in hand-written CUDA kernels reached through the C ABI of
``include/gabo_b200.h`` (``gabotorch_b200/lib/libgabo_b200.so``).  There is no CPU fallback: the first call that
needs the library raises ``GaboError`` when it has not been built or no CUDA device is visible.

Module map (reference module -> here):
    BoManifolds/kernel_utils/kernels_{sphere,spd,nested_spd,nested_sphere}.py -> gabotorch_b200.kernel_utils
    BoManifolds/Riemannian_utils/{sphere,spd}_utils_torch.py   -> gabotorch_b200.riemannian_utils
    BoManifolds/manifold_optimization/manifold_optimize.py     -> gabotorch_b200.manifold_optimization
    BoManifolds/manifold_optimization/manifold_gp_fit.py       -> gabotorch_b200.manifold_gp_fit
    BoManifolds/nested_mappings/nested_spd_utils.py            -> gabotorch_b200.nested_mappings
    BoManifolds/nested_mappings/nested_{spheres,spd}_optimization.py -> gabotorch_b200.nested_optimization
    pymanopt.manifolds.{Sphere,PositiveDefinite}               -> gabotorch_b200.manifolds
"""
__version__ = '0.1.0'

from ._lib import GaboError  # noqa: F401
from .kernel_utils import (SphereGaussianKernel, SphereLaplaceKernel, SpdAffineInvariantGaussianKernel,  # noqa: F401
                           SpdAffineInvariantLaplaceKernel, SpdFrobeniusGaussianKernel,
                           SpdLogEuclideanGaussianKernel, NestedSpdAffineInvariantGaussianKernel,
                           NestedSpdLogEuclideanGaussianKernel, NestedSphereGaussianKernel)
from ._compat import ScaleKernel  # noqa: F401
from .manifolds import Sphere, PositiveDefinite  # noqa: F401
from .manifold_optimization import (ConjugateGradient, TrustRegions, ConstrainedTrustRegions,  # noqa: F401
                                    StrictConstrainedTrustRegions, AugmentedLagrangeMethod,
                                    ExpectedImprovement, ManifoldGP,
                                    gen_batch_initial_conditions_manifold, gen_candidates_manifold,
                                    get_best_candidates, joint_optimize_manifold)
from .nested_mappings import (NestedSpdProjection, NestedSpdReconstruction,  # noqa: F401
                              projection_from_spd_to_nested_spd, projection_from_nested_spd_to_spd,
                              max_eigenvalue_nested_spd_constraint, min_eigenvalue_nested_spd_constraint,
                              random_nested_spd_with_spd_eigenvalue_constraints)
from .gp_fit import fit_gpytorch_model, ExactMarginalLogLikelihood  # noqa: F401
from .manifold_gp_fit import fit_gpytorch_manifold  # noqa: F401
from .nested_optimization import (min_error_reconstruction_cost,  # noqa: F401
                                  optimize_reconstruction_parameters_nested_sphere,
                                  min_affine_invariant_distance_reconstruction_cost,
                                  min_log_euclidean_distance_reconstruction_cost,
                                  optimize_reconstruction_parameters_nested_spd)
from ._compat import GammaPrior  # noqa: F401
