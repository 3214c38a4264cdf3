"""Thin cffi (ABI-mode) bindings of the two in-tree CUDA libraries:

* ``libmonorun_pnp.so``  -- the C ABI declared in include/monorun_pnp.h  (batched uncertainty-PnP solver)
* ``libmonorun_head.so`` -- the C ABI declared in include/monorun_head.h (tcgen05 dense correspondence head)

Both are built in-tree by :func:`build` (``nvcc -gencode arch=compute_100a,code=sm_100a``) and loaded with
``ffi.dlopen``.  There is no CPU fallback: if a library is missing or cannot be loaded the product entry points
raise.
"""
import os
import re
import subprocess
import threading

from cffi import FFI

_PKG = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_PKG)
LIB_PATH = os.environ.get('MRPNP_LIB') or os.path.join(_PKG, 'libmonorun_pnp.so')   # MRPNP_LIB: A/B builds of tools/
HEADER = os.path.join(_ROOT, 'include', 'monorun_pnp.h')
SOURCES = [os.path.join(_PKG, 'csrc', f) for f in ('pnp_capi.cu', 'pnp_kernel.cuh', 'pnp_device.cuh', 'pnp_fast.cuh', 'pnp_kernel_fast.cuh', 'pnp_score.cuh', 'pnp_nms.cuh', 'pnp_noc.cuh', 'pnp_exact_hessian.cuh', 'pnp_6dof.cuh', 'pnp_6dof_fast.cuh', 'lm_dense.cuh')]
HEAD_LIB_PATH = os.environ.get('MRHEAD_LIB') or os.path.join(_PKG, 'libmonorun_head.so')   # MRHEAD_LIB: A/B builds of tools/
HEAD_HEADER = os.path.join(_ROOT, 'include', 'monorun_head.h')
HEAD_SOURCES = [os.path.join(_PKG, 'csrc', f) for f in ('head_capi.cu', 'head_kernels.cuh', 'head_tc.cuh', 'head_carafe_tc.cuh')]

NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-lineinfo', '-std=c++17',
              '-Xcompiler', '-fPIC', '-shared']


def _cdef_from_header(path=None):
    """The cdef is the header itself (minus preprocessor lines), so the binding cannot drift from it."""
    text = open(path or HEADER).read()
    consts = re.findall(r'^#define\s+(MR\w+)\s+\(?(-?\d+)\)?\s*(?:/\*.*)?$', text, re.M)
    body = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    body = '\n'.join(l for l in body.splitlines()
                     if not l.lstrip().startswith('#') and 'extern "C"' not in l and l.strip() != '}')
    for k, v in consts:   # cffi's cdef has no preprocessor: array bounds written with a #define become numbers
        body = re.sub(r'\[\s*%s\s*\]' % k, '[%s]' % v, body)
    return body, {k: int(v) for k, v in consts}


ffi = FFI()
_CDEF, CONST = _cdef_from_header()
ffi.cdef(_CDEF)
_lib = None
_lock = threading.Lock()

head_ffi = FFI()
_HEAD_CDEF, HEAD_CONST = _cdef_from_header(HEAD_HEADER)
head_ffi.cdef(_HEAD_CDEF)
_head_lib = None


def _stale(lib_path, deps):
    if not os.path.exists(lib_path):
        return True
    t = os.path.getmtime(lib_path)
    return any(os.path.getmtime(s) > t for s in deps)


def needs_build():
    return _stale(LIB_PATH, SOURCES + [HEADER]) or _stale(HEAD_LIB_PATH, HEAD_SOURCES + [HEAD_HEADER])


def build(force=False, verbose=False):
    """Compile both CUDA libraries for sm_100a in-tree (nvcc cross-compiles without a GPU)."""
    nvcc = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
    for lib_path, srcs, hdr in ((LIB_PATH, SOURCES, HEADER), (HEAD_LIB_PATH, HEAD_SOURCES, HEAD_HEADER)):
        if not force and not _stale(lib_path, srcs + [hdr]):
            continue
        cmd = [nvcc] + NVCC_FLAGS + ['-ccbin', '/usr/bin/g++', '-I', os.path.join(_ROOT, 'include'), srcs[0], '-o', lib_path]
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


def head_lib():
    """dlopen libmonorun_head.so; raises if it is absent (the tcgen05 head has no fallback inside this call)."""
    global _head_lib
    with _lock:
        if _head_lib is None:
            if not os.path.exists(HEAD_LIB_PATH):
                raise RuntimeError(f'{HEAD_LIB_PATH} is missing: run `python -c "import __graft_entry__ as g; g.build()"`')
            _head_lib = head_ffi.dlopen(HEAD_LIB_PATH)
            if _head_lib.mrhead_version() != HEAD_CONST['MRHEAD_VERSION']:
                raise RuntimeError('libmonorun_head.so does not match include/monorun_head.h; rebuild')
    return _head_lib


def head_check(rc):
    if rc != 0:
        raise RuntimeError(f'libmonorun_head error {rc}: ' + head_ffi.string(head_lib().mrhead_last_error()).decode())


def last_error():
    return ffi.string(lib().mrpnp_last_error()).decode()


def check(rc):
    if rc != 0:
        raise RuntimeError(f'libmonorun_pnp error {rc}: {last_error()}')


EXPORTED = ['mrpnp_default_params', 'mrpnp_create', 'mrpnp_destroy', 'mrpnp_solve', 'mrpnp_solve_dense', 'mrpnp_solve_host',
            'mrpnp_launch_count', 'mrpnp_handed_back_count', 'mrpnp_gather_wait', 'mrpnp_gather_timeouts', 'mrpnp_kernel_info', 'mrpnp_version', 'mrpnp_last_error', 'mrpnp_pose_features',
            'mrpnp_finish_scores', 'mrpnp_score_stage', 'mrpnp_nms_bev', 'mrpnp_solve_noc', 'mrpnp_exact_hessian', 'mrpnp_solve_6dof', 'pnp_uncert']
HEAD_EXPORTED = ['mrhead_create', 'mrhead_destroy', 'mrhead_version', 'mrhead_last_error', 'mrhead_launch_count',
                 'mrhead_pack_input', 'mrhead_conv', 'mrhead_latent_bias', 'mrhead_carafe', 'mrhead_workspace_bytes',
                 'mrhead_forward']
