"""ctypes binding of libgabo_b200.so (the C ABI declared in include/gabo_b200.h).

There is no CPU fallback: importing the product path without the built library raises.
"""
import ctypes
import os

from . import build as _build

_LIB = None

c_i32 = ctypes.c_int
c_i64 = ctypes.c_int64
c_f64 = ctypes.c_double
c_ptr = ctypes.c_void_p

GABO_F32, GABO_F64 = 0, 1
KIND_GAUSS, KIND_LAPLACE, KIND_DIST = 0, 1, 2
SPHERE, SPD = 0, 1
OP_PROJ, OP_RETR, OP_EXP, OP_LOG, OP_TRANSP, OP_PTRANSP, OP_EGRAD2RGRAD = range(7)
MAX_SPHERE_DIM, MAX_SPD_DIM, MAX_TRAIN = 128, 8, 128


class GpDesc(ctypes.Structure):
    _fields_ = [('manifold', ctypes.c_int32), ('dim', ctypes.c_int32), ('n_train', ctypes.c_int32),
                ('compute', ctypes.c_int32), ('x_train', c_ptr), ('alpha', c_ptr), ('minv', c_ptr),
                ('mean', c_f64), ('outputscale', c_f64), ('beta', c_f64), ('best_f', c_f64), ('kxx', c_f64)]


class RcgOpts(ctypes.Structure):
    _fields_ = [('maxiter', ctypes.c_int32), ('ls_maxiter', ctypes.c_int32), ('mingradnorm', c_f64),
                ('minstepsize', c_f64), ('contraction', c_f64), ('suff_decr', c_f64), ('initial_stepsize', c_f64)]


class RtrOpts(ctypes.Structure):
    _fields_ = [('maxiter', ctypes.c_int32), ('mininner', ctypes.c_int32), ('maxinner', ctypes.c_int32),
                ('reserved', ctypes.c_int32), ('mingradnorm', c_f64), ('kappa', c_f64), ('theta', c_f64),
                ('rho_prime', c_f64), ('rho_regularization', c_f64), ('delta_bar', c_f64), ('delta0', c_f64)]


CONS_MAX_EIG, CONS_MIN_EIG = 0, 1


class CtrOpts(ctypes.Structure):
    _fields_ = [('tr', RtrOpts), ('n_constraints', ctypes.c_int32), ('strict', ctypes.c_int32),
                ('kind', ctypes.c_int32 * 2), ('bound', c_f64 * 2), ('delta_cons', c_f64)]


# name -> (restype, argtypes); must list every symbol include/gabo_b200.h declares (tests/test_abi.py checks it)
SIGNATURES = {
    'gabo_version': (c_i32, []),
    'gabo_last_error': (ctypes.c_char_p, []),
    'gabo_sphere_gram': (c_i32, [c_ptr, c_i64, c_ptr, c_i64, c_i32, c_f64, c_i32, c_ptr, c_i32, c_i64, c_ptr]),
    'gabo_sphere_gram_diag': (c_i32, [c_ptr, c_ptr, c_i64, c_i32, c_f64, c_i32, c_ptr, c_i32, c_ptr]),
    'gabo_mandel_unpack': (c_i32, [c_ptr, c_i64, c_i32, c_ptr, c_ptr]),
    'gabo_mandel_pack': (c_i32, [c_ptr, c_i64, c_i32, c_ptr, c_ptr]),
    'gabo_spd_factor_stride': (c_i64, [c_i32]),
    'gabo_spd_factor': (c_i32, [c_ptr, c_i64, c_i32, c_i32, c_ptr, c_ptr, c_ptr]),
    'gabo_spd_factor2': (c_i32, [c_ptr, c_i64, c_ptr, c_i64, c_i32, c_i32, c_ptr, c_ptr, c_ptr, c_ptr]),
    'gabo_spd_ai_gram': (c_i32, [c_ptr, c_i64, c_ptr, c_i64, c_i32, c_f64, c_i32, c_i32, c_i32, c_ptr, c_i32, c_i64,
                                 c_ptr]),
    'gabo_spd_ai_gram_backward': (c_i32, [c_ptr, c_i64, c_ptr, c_i64, c_i32, c_ptr, c_i64, c_i32, c_i32, c_ptr, c_ptr]),
    'gabo_frobenius_gram': (c_i32, [c_ptr, c_i64, c_ptr, c_i64, c_i32, c_f64, c_i32, c_ptr, c_i32, c_i64, c_ptr]),
    'gabo_spd_logm': (c_i32, [c_ptr, c_i64, c_i32, c_ptr, c_ptr]),
    'gabo_sym_eig': (c_i32, [c_ptr, c_i64, c_i32, c_ptr, c_ptr, c_ptr, c_ptr]),
    'gabo_weighted_points_sum': (c_i32, [c_ptr, c_ptr, c_i64, c_i64, c_i64, c_i32, c_i32, c_ptr, c_i32, c_ptr, c_ptr]),
    'gabo_spd_logm_backward': (c_i32, [c_ptr, c_ptr, c_i64, c_i32, c_ptr, c_ptr]),
    'gabo_nested_spd_project_backward': (c_i32, [c_ptr, c_ptr, c_ptr, c_i64, c_i32, c_i32, c_ptr, c_ptr, c_ptr]),
    'gabo_sphere_op': (c_i32, [c_i32, c_ptr, c_ptr, c_ptr, c_i64, c_i32, c_ptr, c_ptr]),
    'gabo_sphere_dist': (c_i32, [c_ptr, c_ptr, c_i64, c_i32, c_ptr, c_ptr]),
    'gabo_nested_sphere_project': (c_i32, [c_ptr, c_i64, c_i32, c_i32, c_ptr, c_ptr, c_ptr, c_ptr]),
    'gabo_spd_op': (c_i32, [c_i32, c_ptr, c_ptr, c_ptr, c_i64, c_i32, c_ptr, c_ptr]),
    'gabo_spd_scalar': (c_i32, [c_i32, c_ptr, c_ptr, c_ptr, c_i64, c_i32, c_ptr, c_ptr]),
    'gabo_ei_eval': (c_i32, [ctypes.POINTER(GpDesc), c_ptr, c_i64, c_ptr, c_ptr, c_ptr]),
    'gabo_acq_rcg': (c_i32, [ctypes.POINTER(GpDesc), c_ptr, c_i64, ctypes.POINTER(RcgOpts), c_ptr, c_ptr, c_ptr,
                             c_ptr]),
    'gabo_acq_rtr': (c_i32, [ctypes.POINTER(GpDesc), c_ptr, c_i64, ctypes.POINTER(RtrOpts), c_ptr, c_ptr, c_ptr,
                             c_ptr]),
    'gabo_acq_ctr': (c_i32, [ctypes.POINTER(GpDesc), c_ptr, c_i64, ctypes.POINTER(CtrOpts), c_ptr, c_ptr, c_ptr,
                             c_ptr]),
    'gabo_argmax_records': (c_i32, [c_ptr, c_ptr, c_i64, c_ptr, c_ptr, c_ptr]),
    'gabo_nested_spd_project_f64': (c_i32, [c_ptr, c_i64, c_i32, c_i32, c_ptr, c_ptr, c_ptr]),
    'gabo_nested_projection_pack_size': (c_i64, [c_i32, c_i32]),
    'gabo_nested_projection_matrix': (c_i32, [c_ptr, c_i32, c_i32, c_ptr, c_ptr]),
    'gabo_nested_spd_project': (c_i32, [c_ptr, c_i64, c_i32, c_i32, c_ptr, c_ptr, c_ptr]),
    'gabo_nested_sphere_chain': (c_i32, [c_ptr, c_i64, c_i32, c_i32, c_ptr, c_ptr, c_ptr, c_ptr]),
    'gabo_nested_sphere_to_nested': (c_i32, [c_ptr, c_i64, c_i32, c_ptr, c_f64, c_ptr, c_ptr]),
    'gabo_nested_sphere_reconstruct': (c_i32, [c_ptr, c_i64, c_i32, c_i32, c_ptr, c_ptr, c_ptr, c_ptr]),
    'gabo_spd_sqrtm': (c_i32, [c_ptr, c_i64, c_i32, c_ptr, c_ptr]),
    'gabo_nested_spd_reconstruct_pack_size': (c_i64, [c_i32, c_i32]),
    'gabo_nested_spd_reconstruct_setup': (c_i32, [c_ptr, c_ptr, c_ptr, c_ptr, c_i32, c_i32, c_ptr, c_ptr, c_ptr]),
    'gabo_nested_spd_reconstruct': (c_i32, [c_ptr, c_ptr, c_i64, c_i32, c_i32, c_ptr, c_ptr, c_ptr]),
    'gabo_gp_mll': (c_i32, [c_ptr, c_i64, c_ptr, c_ptr, c_i64, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr]),
    'gabo_gp_factor': (c_i32, [c_ptr, c_i64, c_ptr, c_f64, c_f64, c_f64, c_ptr, c_ptr, c_ptr, c_ptr]),
    'gabo_gp_fit': (c_i32, [c_ptr, c_i64, c_ptr, c_ptr, c_i64, c_f64, c_f64, c_ptr, c_ptr, c_i32, c_f64, c_f64, c_ptr,
                            c_ptr, c_ptr, c_ptr]),
}


class GaboError(RuntimeError):
    pass


def lib_path():
    return _build.lib_path()


def load():
    """Load (once) and return the ctypes handle.  Raises if the library has not been built."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = lib_path()
    if not os.path.exists(path):
        raise GaboError('%s is missing: run `python -m gabotorch_b200.build` (or __graft_entry__.build()). '
                        'gabotorch_b200 has no CPU fallback.' % path)
    lib = ctypes.CDLL(path)
    missing = [name for name in SIGNATURES if not hasattr(lib, name)]
    if missing and os.environ.get('GABO_DEV_PARTIAL'):
        missing_ok = set(missing)
        missing = []
    else:
        missing_ok = set()
    if missing:
        raise GaboError('%s does not export %s: stale build, run `python -m gabotorch_b200.build --force`'
                        % (path, ', '.join(missing)))
    for name, (res, args) in SIGNATURES.items():
        if name in missing_ok:
            continue
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _LIB = lib
    return lib


def check(code, what=''):
    if code != 0:
        msg = load().gabo_last_error()
        raise GaboError('%s failed with code %d: %s' % (what or 'gabo call', code, (msg or b'').decode()))


def stream_ptr():
    """The current torch CUDA stream as a void*."""
    import torch
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
