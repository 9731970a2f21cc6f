"""In-tree build of the C-ABI CUDA library (csrc/*.cu -> lib/libmsda_b200.so) for sm_100a.

Plain `nvcc -shared`: the library has no torch / ATen dependency (include/msda_b200.h is the whole
boundary), so it cross-compiles on a CPU-only box and the built .so travels to the GPU box.
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIBDIR = os.path.join(HERE, 'lib')
LIBNAME = 'libmsda_b200.so'
SOURCES = ['msda_fwd.cu', 'msda_fwd_smem.cu', 'msda_bwd.cu', 'msda_bwd_cell.cu', 'msda_bwd_sorted.cu', 'adapter_dwconv.cu', 'adapter_layernorm.cu', 'adapter_colsum.cu', 'adapter_residual.cu', 'msda_abi.cu']
# (source, extra flags, object name): msda_fwd.cu / msda_bwd.cu are compiled once per value dtype so that the three sets
# of template instantiations build in parallel (MSDA_TU = 0: f32 + f64 + dispatch, 1: bf16, 2: f16)
UNITS = [(s, [], s.replace('.cu', '.o')) for s in SOURCES] + [
    ('msda_fwd.cu', ['-DMSDA_TU=1'], 'msda_fwd_bf16.o'), ('msda_fwd.cu', ['-DMSDA_TU=2'], 'msda_fwd_f16.o'),
    ('msda_bwd.cu', ['-DMSDA_TU=1'], 'msda_bwd_bf16.o'), ('msda_bwd.cu', ['-DMSDA_TU=2'], 'msda_bwd_f16.o'),
    ('msda_bwd_cell.cu', ['-DMSDA_TU=1'], 'msda_bwd_cell_bf16.o'), ('msda_bwd_cell.cu', ['-DMSDA_TU=2'], 'msda_bwd_cell_f16.o'),
    ('msda_bwd_sorted.cu', ['-DMSDA_TU=1'], 'msda_bwd_sorted_bf16.o'), ('msda_bwd_sorted.cu', ['-DMSDA_TU=2'], 'msda_bwd_sorted_f16.o')]
HEADERS = ['msda_common.cuh', 'msda_cell_common.cuh', os.path.join('..', '..', 'include', 'msda_b200.h')]
NVCC_FLAGS = [
    '-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-lineinfo', '-std=c++17',
    '-Xcompiler', '-fPIC',
]


def lib_path():
    return os.path.join(LIBDIR, LIBNAME)


def _nvcc():
    cand = shutil.which('nvcc') or '/usr/local/cuda/bin/nvcc'
    if not os.path.exists(cand):
        raise RuntimeError('nvcc not found: cannot build %s' % LIBNAME)
    return cand


def is_stale():
    out = lib_path()
    if not os.path.exists(out):
        return True
    t = os.path.getmtime(out)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    """Compile every .cu into lib/libmsda_b200.so (parallel per-file objects, then link)."""
    if not force and not is_stale():
        return lib_path()
    os.makedirs(LIBDIR, exist_ok=True)
    objdir = os.path.join(LIBDIR, 'obj')
    os.makedirs(objdir, exist_ok=True)
    nvcc = _nvcc()
    procs = []
    objs = []
    for s, extra, oname in UNITS:
        o = os.path.join(objdir, oname)
        objs.append(o)
        cmd = [nvcc] + NVCC_FLAGS + os.environ.get('MSDA_NVCC_EXTRA', '').split() + extra + (['-Xptxas', '-v'] if verbose else []) + ['-c', '-o', o, os.path.join(CSRC, s)]
        procs.append((cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for cmd, pr in procs:
        out, _ = pr.communicate()
        if verbose and out:
            sys.stderr.write(out)
        if pr.returncode != 0:
            raise RuntimeError('nvcc failed: %s\n%s' % (' '.join(cmd), out))
    tmp = lib_path() + '.tmp'
    link = [nvcc, '-shared', '-cudart', 'static', '-o', tmp] + objs
    r = subprocess.run(link, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError('link failed: %s\n%s' % (' '.join(link), r.stdout))
    os.replace(tmp, lib_path())
    return lib_path()


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
