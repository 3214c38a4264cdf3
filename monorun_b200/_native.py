"""Thin cffi (ABI-mode) binding of ``libmonorun_pnp.so`` -- the C ABI declared in include/monorun_pnp.h.

The library is built in-tree by :func:`build` (``nvcc -gencode arch=compute_100a,code=sm_100a``) and
loaded with ``ffi.dlopen``.  There is no CPU fallback: if the library is missing or cannot be loaded
the import of any product entry point raises.
"""
import os
import re
import subprocess
import threading

from cffi import FFI

_PKG = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_PKG)
LIB_PATH = os.path.join(_PKG, 'libmonorun_pnp.so')
HEADER = os.path.join(_ROOT, 'include', 'monorun_pnp.h')
SOURCES = [os.path.join(_PKG, 'csrc', f) for f in ('pnp_capi.cu', 'pnp_kernel.cuh', 'pnp_device.cuh', 'pnp_kernel_pair.cuh')]

NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-lineinfo', '-std=c++17',
              '-Xcompiler', '-fPIC', '-shared']


def _cdef_from_header():
    """The cdef is the header itself (minus preprocessor lines), so the binding cannot drift from it."""
    text = open(HEADER).read()
    consts = re.findall(r'^#define\s+(MRPNP_\w+)\s+\(?(-?\d+)\)?\s*(?:/\*.*)?$', text, re.M)
    body = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    body = '\n'.join(l for l in body.splitlines()
                     if not l.lstrip().startswith('#') and 'extern "C"' not in l and l.strip() != '}')
    return body, {k: int(v) for k, v in consts}


ffi = FFI()
_CDEF, CONST = _cdef_from_header()
ffi.cdef(_CDEF)
_lib = None
_lock = threading.Lock()


def needs_build():
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    return any(os.path.getmtime(s) > t for s in SOURCES + [HEADER])


def build(force=False, verbose=False):
    """Compile the CUDA extension for sm_100a in-tree (nvcc cross-compiles without a GPU)."""
    if not force and not needs_build():
        return LIB_PATH
    nvcc = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
    cmd = [nvcc] + NVCC_FLAGS + ['-ccbin', '/usr/bin/g++', '-I', os.path.join(_ROOT, 'include'),
                                 SOURCES[0], '-o', LIB_PATH]
    if verbose:
        cmd.insert(1, '-Xptxas=-v')
    subprocess.check_call(cmd)
    return LIB_PATH


def lib():
    """dlopen the extension; raises if it is absent (no fallback path exists)."""
    global _lib
    with _lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                raise RuntimeError(
                    f'{LIB_PATH} is missing: run `python -c "import __graft_entry__ as g; g.build()"`; '
                    'monorun_b200 has no CPU or PyTorch fallback for the PnP solver')
            _lib = ffi.dlopen(LIB_PATH)
            if _lib.mrpnp_version() != CONST['MRPNP_VERSION']:
                raise RuntimeError('libmonorun_pnp.so does not match include/monorun_pnp.h; rebuild')
    return _lib


def last_error():
    return ffi.string(lib().mrpnp_last_error()).decode()


def check(rc):
    if rc != 0:
        raise RuntimeError(f'libmonorun_pnp error {rc}: {last_error()}')


EXPORTED = ['mrpnp_default_params', 'mrpnp_create', 'mrpnp_destroy', 'mrpnp_solve', 'mrpnp_solve_dense', 'mrpnp_solve_host',
            'mrpnp_launch_count', 'mrpnp_kernel_info', 'mrpnp_version', 'mrpnp_last_error']
