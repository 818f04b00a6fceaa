"""Build libfeabas_cuda.so in-tree for sm_100a (nvcc cross-compiles without a GPU).

    python -m feabas_b200.csrc.build [--force] [-v]

Every translation unit is compiled to an object file (only when it or a header is newer) and the
objects are linked into one shared library next to the sources.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, 'libfeabas_cuda.so')
OBJ_DIR = os.path.join(HERE, 'build')
HEADER = os.path.join('..', '..', 'include', 'feabas_cuda.h')
FAST_DEPS = ['fb_fast_tu.inc', 'fb_fast_groups.h', 'fb_xcorr_fast.cuh', 'fb_xcorr.cuh', 'fb_regfft.cuh', 'fb_fft.cuh', 'fb_gfft.cuh', HEADER]
WF_DEPS = ['fb_wf_tu.inc', 'fb_wf_groups.h', 'fb_xcorr_wf.cuh', 'fb_xcorr.cuh', 'fb_regfft.cuh', 'fb_fft.cuh', 'fb_gfft.cuh', HEADER]
# translation unit -> headers it includes
UNITS = {
    'fb_xcorr.cu': ['fb_xcorr.cuh', 'fb_xcorr_fast.cuh', 'fb_fast_groups.h', 'fb_wf_groups.h', 'fb_xcorr_wf.cuh', 'fb_regfft.cuh', 'fb_fft.cuh', 'fb_gfft.cuh', 'fb_host_plan.h', 'fb_common.h', HEADER],
    # the register-resident fast path, one translation unit per group of line lengths (parallel build)
    'fb_fast_pow2.cu': FAST_DEPS, 'fb_fast_big.cu': FAST_DEPS, 'fb_fast_r3.cu': FAST_DEPS, 'fb_fast_r5.cu': FAST_DEPS, 'fb_fast_r5b.cu': FAST_DEPS,
    'fb_image.cu': ['fb_common.h', HEADER],
    'fb_image_ext.cu': ['fb_common.h', HEADER],
    # the warp-fused kernel (whole pair on one SM, register transforms), one translation unit per group of grids
    'fb_wf_a.cu': WF_DEPS, 'fb_wf_b.cu': WF_DEPS, 'fb_wf_c.cu': WF_DEPS,
}
SOURCES = list(UNITS)
NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17',
              '-Xcompiler', '-fPIC', '-diag-suppress', '68'] + os.environ.get('FEABAS_NVCC_FLAGS', '').split()


def _nvcc():
    for cand in (os.environ.get('NVCC'), '/usr/local/cuda/bin/nvcc', 'nvcc'):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return 'nvcc'


def _mtime(name):
    path = os.path.join(HERE, name)
    return os.path.getmtime(path) if os.path.exists(path) else 0.0


def _obj(unit):
    return os.path.join(OBJ_DIR, os.path.splitext(unit)[0] + '.o')


def _unit_stale(unit):
    obj = _obj(unit)
    if not os.path.exists(obj):
        return True
    t = os.path.getmtime(obj)
    return any(_mtime(d) > t for d in [unit] + UNITS[unit])


def stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(_unit_stale(u) or os.path.getmtime(_obj(u)) > t for u in UNITS)


def _run(cmd, verbose):
    res = subprocess.run(cmd, cwd=HERE, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError('nvcc failed:\n' + ' '.join(cmd) + '\n' + res.stdout + res.stderr)
    if verbose:
        print(res.stderr)


def build(force=False, verbose=False):
    if not force and not stale():
        return LIB
    os.makedirs(OBJ_DIR, exist_ok=True)
    procs = []
    for unit in UNITS:
        if force or _unit_stale(unit):
            cmd = [_nvcc()] + NVCC_FLAGS + (['-Xptxas', '-v'] if verbose else []) + ['-c', unit, '-o', _obj(unit)]
            procs.append((cmd, subprocess.Popen(cmd, cwd=HERE, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for cmd, proc in procs:
        out, _ = proc.communicate()
        if proc.returncode != 0:
            raise RuntimeError('nvcc failed:\n' + ' '.join(cmd) + '\n' + out)
        if verbose:
            print(out)
    _run([_nvcc(), '-shared', '-o', LIB] + [_obj(u) for u in UNITS], verbose)
    return LIB


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
