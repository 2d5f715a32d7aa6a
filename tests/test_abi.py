"""The C-ABI library loads and exports every symbol include/gabo_b200.h declares; argument validation returns error
codes without touching the GPU.  CPU only (no compute calls)."""
import ctypes
import os
import re

import pytest

from gabotorch_b200 import _lib, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, 'include', 'gabo_b200.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(gabo_[a-z0-9_]+)\s*\(', text)))


def test_library_is_built_in_tree():
    assert os.path.exists(build.lib_path()), 'run python -m gabotorch_b200.build'
    assert os.path.dirname(build.lib_path()).startswith(ROOT)


def test_every_header_symbol_is_exported_and_bound(lib):
    syms = header_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(lib, s), 'libgabo_b200.so does not export %s' % s
        assert s in _lib.SIGNATURES, 'no ctypes signature for %s' % s
    assert sorted(_lib.SIGNATURES) == syms


def test_version_and_struct_layout(lib):
    assert lib.gabo_version() == 100
    assert ctypes.sizeof(_lib.GpDesc) == 16 + 3 * 8 + 5 * 8
    assert ctypes.sizeof(_lib.RcgOpts) == 8 + 5 * 8
    assert lib.gabo_spd_factor_stride(3) == 12 and lib.gabo_spd_factor_stride(8) == 72
    assert lib.gabo_spd_factor_stride(9) == -1
    # 27 k-steps x 32 lanes x one 16-row operator tile x (hi, lo) float4 + the canonical (hi, lo) image of the tcgen05 kernel
    assert lib.gabo_nested_projection_pack_size(20, 5) == 27 * 32 * 8 + 2 * 16 * 216
    assert lib.gabo_nested_projection_pack_size(5, 9) == -1


def test_argument_errors_are_codes_not_crashes(lib):
    null = ctypes.c_void_p(None)
    fake = ctypes.c_void_p(4096)  # aligned, never dereferenced: validation fails first
    assert lib.gabo_sphere_gram(null, 4, null, 4, 3, 1.0, 0, null, 0, 4, null) == -1
    assert b'null' in lib.gabo_last_error()
    assert lib.gabo_sphere_gram(fake, -1, fake, 4, 3, 1.0, 0, fake, 0, 4, null) == -1
    assert lib.gabo_sphere_gram(fake, 4, fake, 4, 3, 1.0, 7, fake, 0, 4, null) == -1       # bad kind
    assert lib.gabo_sphere_gram(fake, 4, fake, 4, 3, 1.0, 0, fake, 0, 2, null) == -1       # ld_out < n2
    assert lib.gabo_sphere_gram(ctypes.c_void_p(4100), 4, fake, 4, 3, 1.0, 0, fake, 0, 4, null) == -2  # alignment
    assert lib.gabo_sphere_gram(fake, 0, fake, 4, 3, 1.0, 0, fake, 0, 4, null) == 0        # empty input: nothing to do
    assert lib.gabo_spd_ai_gram(fake, 4, fake, 4, 9, 1.0, 0, 0, 0, fake, 0, 4, null) == -1  # d > 8
    assert b'outside' in lib.gabo_last_error()
    assert lib.gabo_spd_ai_gram(fake, 4, fake, 5, 3, 1.0, 0, 0, 1, fake, 0, 5, null) == -1  # symmetric needs same set
    assert lib.gabo_mandel_unpack(null, 3, 3, null, null) == -1
    assert lib.gabo_spd_logm(fake, 3, 12, fake, null) == -1
    assert lib.gabo_sym_eig(fake, 3, 33, fake, null, null, null) == -1
    assert lib.gabo_sym_eig(fake, 0, 33, fake, null, null, null) == 0
    assert lib.gabo_argmax_records(null, null, 4, null, null, null) == -1
    desc = _lib.GpDesc(0, 3, 0, 0, 4096, 4096, 4096, 0.0, 1.0, 1.0, 0.0, 1.0)
    assert lib.gabo_ei_eval(ctypes.byref(desc), fake, 4, fake, null, null) == -1           # n_train = 0
    desc.n_train = 4
    desc.manifold = 5
    assert lib.gabo_ei_eval(ctypes.byref(desc), fake, 4, fake, null, null) == -1
    assert lib.gabo_nested_spd_project(fake, 8, 4, 5, fake, fake, null) == -1              # d > D
    assert lib.gabo_nested_sphere_chain(fake, 4, 70, 3, fake, fake, fake, null) == -1         # D > 64
    assert lib.gabo_nested_sphere_chain(null, 4, 5, 3, null, null, null, null) == -1
    assert lib.gabo_nested_sphere_chain(null, 4, 5, 5, null, null, null, null) == 0           # no level: nothing to do
    assert lib.gabo_nested_sphere_reconstruct(fake, 4, 6, 5, fake, fake, fake, null) == -1    # d_latent > D
    assert lib.gabo_nested_sphere_to_nested(fake, 4, 1, fake, 1.0, fake, null) == -1
    assert lib.gabo_spd_sqrtm(fake, 3, 9, fake, null) == -1
    assert lib.gabo_nested_spd_reconstruct_pack_size(20, 5) == 2 * 20 * 5 + 400 + (25 + 15) * 400
    assert lib.gabo_nested_spd_reconstruct_pack_size(5, 5) == 0
    assert lib.gabo_nested_spd_reconstruct(fake, fake, 4, 40, 5, fake, fake, null) == -1      # D > 32
    assert lib.gabo_nested_spd_reconstruct_setup(fake, fake, fake, fake, 5, 5, fake, fake, null) == -1
    assert lib.gabo_gp_mll(fake, 129, fake, fake, 1, fake, null, null, null, fake, null) == -1   # n > 128
    assert lib.gabo_gp_mll(null, 8, null, null, 1, null, null, null, null, null, null) == -1
    assert lib.gabo_gp_mll(fake, 8, fake, fake, 0, fake, null, null, null, fake, null) == 0      # empty batch
    assert lib.gabo_gp_factor(fake, 200, fake, 1.0, 1.0, 0.0, fake, fake, fake, null) == -1     # n > 128
    assert lib.gabo_gp_factor(fake, 8, fake, 1.0, 1.0, 0.0, null, fake, fake, null) == -1
    opts = _lib.RtrOpts(10, 1, 0, 0, 1e-6, 0.1, 1.0, 0.1, 1e3, 0.0, 0.0)
    desc.manifold, desc.dim, desc.n_train = 1, 3, 4                                             # SPD(3)
    copts = _lib.CtrOpts()
    copts.tr, copts.delta_cons, copts.n_constraints = opts, 1e-6, 3
    assert lib.gabo_acq_ctr(ctypes.byref(desc), fake, 4, ctypes.byref(copts), fake, null, null, null) == -1   # > 2 constraints
    copts.n_constraints, copts.kind[0] = 1, 7
    assert lib.gabo_acq_ctr(ctypes.byref(desc), fake, 4, ctypes.byref(copts), fake, null, null, null) == -1   # unknown kind
    copts.kind[0], copts.delta_cons = 0, 0.0
    assert lib.gabo_acq_ctr(ctypes.byref(desc), fake, 4, ctypes.byref(copts), fake, null, null, null) == -1
    copts.delta_cons = 1e-6
    assert lib.gabo_acq_ctr(ctypes.byref(desc), fake, 0, ctypes.byref(copts), fake, null, null, null) == 0    # no restart
    desc.manifold = 0
    assert lib.gabo_acq_ctr(ctypes.byref(desc), fake, 4, ctypes.byref(copts), fake, null, null, null) == -4   # sphere
    desc.manifold, desc.dim = 0, 20                                                              # beyond the register kernel
    assert lib.gabo_acq_rtr(ctypes.byref(desc), fake, 4, ctypes.byref(opts), fake, null, null, null) == -4
    opts.rho_prime = 0.5
    assert lib.gabo_acq_rtr(ctypes.byref(desc), fake, 4, ctypes.byref(opts), fake, null, null, null) == -1


def test_product_path_has_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    import gabotorch_b200 as g
    k = g.SphereGaussianKernel(beta_min=6.5)
    x = torch.nn.functional.normalize(torch.randn(5, 3, dtype=torch.float64), dim=-1)
    with torch.no_grad(), pytest.raises(g.GaboError):
        k.forward(x, x)
    with pytest.raises(g.GaboError):
        g.Sphere(3).exp(x[0].numpy(), 0.1 * x[1].numpy())


def test_product_code_never_imports_the_oracle():
    pkg = os.path.join(ROOT, 'gabotorch_b200')
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.h')):
                text = open(os.path.join(dirpath, f)).read()
                assert 'import oracle' not in text and 'from oracle' not in text, f


def test_header_is_plain_c_and_struct_layouts_match_the_ctypes_mirrors(tmp_path):
    # include/gabo_b200.h is the drop-in boundary: it must compile as C (no C++ / CUDA types in the signatures) and
    # the structs a binding passes by pointer must have exactly the layout gabotorch_b200/_lib.py mirrors
    import shutil
    import subprocess
    gcc = shutil.which('gcc')
    if gcc is None:
        pytest.skip('gcc not available')
    fields = {'gabo_gp_desc': _lib.GpDesc, 'gabo_rcg_opts': _lib.RcgOpts, 'gabo_rtr_opts': _lib.RtrOpts,
              'gabo_ctr_opts': _lib.CtrOpts}
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "gabo_b200.h"', 'int main(void) {']
    for cname, mirror in fields.items():
        lines.append('  printf("%s %%zu\\n", sizeof(%s));' % (cname, cname))
        for fname, _ in mirror._fields_:
            lines.append('  printf("%s.%s %%zu\\n", offsetof(%s, %s));' % (cname, fname, cname, fname))
    lines += ['  return 0;', '}']
    src = tmp_path / 'layout.c'
    src.write_text('\n'.join(lines))
    exe = tmp_path / 'layout'
    subprocess.run([gcc, '-std=c99', '-Wall', '-Werror', '-pedantic', '-I', os.path.join(ROOT, 'include'), str(src),
                    '-o', str(exe)], check=True, capture_output=True)
    out = dict(l.split() for l in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.splitlines())
    for cname, mirror in fields.items():
        assert int(out[cname]) == ctypes.sizeof(mirror), cname
        for fname, _ in mirror._fields_:
            assert int(out['%s.%s' % (cname, fname)]) == getattr(mirror, fname).offset, '%s.%s' % (cname, fname)
    assert ctypes.sizeof(_lib.RtrOpts) == 16 + 7 * 8
