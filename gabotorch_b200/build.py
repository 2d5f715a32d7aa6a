"""Build libgabo_b200.so (sm_100a only) in-tree with nvcc.

    python -m gabotorch_b200.build [--force]

One object per .cu under ``csrc/`` (compiled in parallel), linked into ``gabotorch_b200/lib/libgabo_b200.so`` with a
static CUDA runtime, so the library has no link-time dependency on torch.  The .so is git-ignored but travels to the
GPU box with the tree.
"""
import concurrent.futures
import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIBDIR = os.path.join(HERE, 'lib')
OBJDIR = os.path.join(HERE, 'build')
LIBNAME = 'libgabo_b200.so'

NVCC_FLAGS = [
    '-gencode', 'arch=compute_100a,code=sm_100a',
    '-O3', '-std=c++17', '-lineinfo',
    '-Xcompiler', '-fPIC,-O3',
    '--expt-relaxed-constexpr',
    '-Xptxas', '-v',
]


def _nvcc():
    for cand in (os.environ.get('NVCC'), shutil.which('nvcc'), '/usr/local/cuda/bin/nvcc'):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError('nvcc not found (set NVCC or CUDA_HOME); gabotorch_b200 has no CPU fallback')


def _sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith('.cu'))


def _headers():
    hs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith('.cuh')]
    hs.append(os.path.join(os.path.dirname(HERE), 'include', 'gabo_b200.h'))
    return sorted(hs)


def _digest(paths, extra=''):
    h = hashlib.sha256(extra.encode())
    for p in paths:
        with open(p, 'rb') as f:
            h.update(f.read())
    return h.hexdigest()


def _compile_one(nvcc, src, hdr_digest):
    obj = os.path.join(OBJDIR, os.path.basename(src)[:-3] + '.o')
    stamp = obj + '.sha'
    want = _digest([src], hdr_digest + ' '.join(NVCC_FLAGS))
    if os.path.exists(obj) and os.path.exists(stamp) and open(stamp).read() == want:
        return obj, '', False
    cmd = [nvcc] + NVCC_FLAGS + ['-c', src, '-o', obj]
    p = subprocess.run(cmd, capture_output=True, text=True)
    if p.returncode != 0:
        raise RuntimeError('nvcc failed for %s:\n%s\n%s' % (src, p.stdout, p.stderr))
    with open(stamp, 'w') as f:
        f.write(want)
    return obj, p.stderr, True


def lib_path():
    return os.path.join(LIBDIR, LIBNAME)


def build(force=False, verbose=False):
    """Compile every CUDA source for sm_100a and link the shared library.  Returns its path."""
    os.makedirs(LIBDIR, exist_ok=True)
    os.makedirs(OBJDIR, exist_ok=True)
    if force:
        for f in os.listdir(OBJDIR):
            os.remove(os.path.join(OBJDIR, f))
    nvcc = _nvcc()
    hdr_digest = _digest(_headers())
    srcs = _sources()
    objs, rebuilt, log = [], False, []
    with concurrent.futures.ThreadPoolExecutor(max_workers=min(len(srcs), os.cpu_count() or 4)) as ex:
        for obj, err, did in ex.map(lambda s: _compile_one(nvcc, s, hdr_digest), srcs):
            objs.append(obj)
            rebuilt = rebuilt or did
            log.append(err)
    out = lib_path()
    if rebuilt or not os.path.exists(out):
        cmd = [nvcc, '-shared', '-gencode', 'arch=compute_100a,code=sm_100a', '-Xcompiler', '-fPIC',
               '-cudart', 'static', '-o', out] + objs
        p = subprocess.run(cmd, capture_output=True, text=True)
        if p.returncode != 0:
            raise RuntimeError('link failed:\n%s\n%s' % (p.stdout, p.stderr))
    if verbose:
        sys.stderr.write('\n'.join(log))
    with open(os.path.join(OBJDIR, 'ptxas.log'), 'w') as f:
        f.write('\n'.join(log))
    return out


if __name__ == '__main__':
    path = build(force='--force' in sys.argv, verbose='-v' in sys.argv)
    print(path)
