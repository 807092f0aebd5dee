"""ctypes binding of libcffm_b200.so (the C ABI declared in include/cffm_b200.h).

This is the ONLY way arithmetic of the hot path is executed: there is no PyTorch / CPU
fallback.  If the shared library is missing or the current device is not an sm_100 part,
every call raises -- loudly -- instead of degrading to another implementation.
"""
import ctypes
import os
import re
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libcffm_b200.so")
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "cffm_b200.h")

ACT_NONE, ACT_GELU, ACT_RELU = 0, 1, 2
GEMM_TCGEN05, GEMM_CHECK = 0, 1

_STATUS = {1: "CFFM_E_BADARG", 2: "CFFM_E_UNSUPPORTED", 3: "CFFM_E_ARCH", 4: "CFFM_E_DRIVER"}

vp, i32, i64, f32 = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_float

# argument types of every entry point, in header order (checked against the header by
# tests/test_abi_exports.py so the two cannot drift apart)
SIGNATURES = {
    "cffm_abi_version": ([], i32),
    "cffm_last_error": ([], ctypes.c_char_p),
    "cffm_device_check": ([], i32),
    "cffm_current_device": ([], i32),
    "cffm_gemm_f16": ([vp, i64, vp, i64, vp, vp, i64, vp, i64, vp, i64, i32, i32, i32, i32, i32, vp], i32),
    "cffm_gemm_f16_ln": ([vp, i64, vp, i64, vp, vp, i64, vp, i64, vp, vp, f32, vp, i64, i32, i32, i32, vp], i32),
    "cffm_gemm_f16_ln_chain": ([vp, i64, vp, i64, vp, vp, i64, vp, vp, f32, vp, vp, f32, vp, i64, i32, i32, i32, vp], i32),
    "cffm_gemm_f16_splitk": ([vp, i64, vp, i64, vp, i32, i32, i32, i32, vp], i32),
    "cffm_splitk_plan": ([i32, i32, i32], i32),
    "cffm_conv_gemm_f16_ln": ([vp, i32, i32, i32, i32, i32, i32, i32, vp, i64, vp, vp, i64, vp, vp, f32, vp, vp, f32, vp, i64, i32, vp], i32),
    "cffm_conv_gemm_f16_splitk": ([vp, i32, i32, i32, i32, i32, i32, i32, vp, i64, vp, i32, i32, vp], i32),
    "cffm_patch_embed_s1_supported": ([i32, i32, i32, i32, i32, i32], i32),
    "cffm_patch_embed_s1": ([vp, i32, i32, i32, vp, vp, vp, vp, f32, vp, vp, f32, vp, vp, i32, vp], i32),
    "cffm_mixffn_tail_supported": ([i32, i32], i32),
    "cffm_mixffn_tail": ([vp, i32, i32, i32, i32, vp, vp, vp, i64, vp, vp, vp, vp, vp, f32, vp, i32, vp], i32),
    "cffm_layernorm_sum": ([vp, i32, vp, vp, vp, f32, vp, i64, vp, i64, i32, i32, vp], i32),
    "cffm_layernorm_chain": ([vp, i32, vp, vp, vp, f32, vp, i64, vp, vp, f32, vp, i64, i32, i32, vp], i32),
    "cffm_layernorm": ([vp, i32, i64, vp, vp, f32, vp, i64, vp, i64, i32, i32, vp], i32),
    "cffm_im2col": ([vp, i32, i32, i32, i32, i32, i32, i32, i32, vp, i32, vp], i32),
    "cffm_mha_f16": ([vp, i64, vp, vp, i64, vp, i64, i32, i32, i32, i32, i32, f32, vp], i32),
    "cffm_dwconv3x3_gelu": ([vp, vp, vp, vp, i32, i32, i32, i32, vp], i32),
    "cffm_head_fuse": ([vp, vp, vp, vp, i32, i32, i32, i32, i32, i32, i32, i32, i32, i32, i32, vp, vp, vp, i64, vp,
                        i64, vp], i32),
    "cffm_cffa_norm": ([vp, vp, vp, f32, vp, vp, i32, i32, i32, i32, i32, i32, i32, vp], i32),
    "cffm_cffa_norm_frames": ([vp, vp, vp, f32, vp, vp, i32, i32, i32, i32, i32, i32, i32, vp], i32),
    "cffm_cffa_pool": ([vp, i32, i32, i32, i32, i32, vp, vp, vp, vp], i32),
    "cffm_cffa_pool_part": ([vp, i32, i32, i32, i32, i32, vp, vp, vp, vp], i32),
    "cffm_cffa_pool_level": ([vp, i32, i32, i32, i32, i32, vp, vp, vp, vp], i32),
    "cffm_resize_u8": ([vp, i32, i32, i32, vp, i32, i32, vp], i32),
    "cffm_resize_normalize_u8": ([vp, i32, i32, i32, vp, i64, i32, i32, vp, vp, i32, vp], i32),
    "cffm_kmeans_prepare": ([vp, vp, vp, vp, i32, i32, i32, vp], i32),
    "cffm_kmeans_assign": ([vp, i64, vp, i32, i32, i32, i32, vp, vp, vp, vp], i32),
    "cffm_kmeans_update": ([vp, i32, vp, vp, vp, vp, vp, vp, vp, vp, i32, i32, i32, vp], i32),
    "cffm_transpose_f16": ([vp, i32, i32, vp, i32, vp], i32),
    "cffm_cfm_attention": ([vp, vp, vp, vp, i32, i32, i32, i32, i32, f32, vp], i32),
    "cffm_cfm_attention_slots": ([vp, vp, i64, vp, vp, vp, i64, i64, i64, i32, i32, i32, i64, i32, vp, vp, vp, i32, i32, i32, i32, i32, f32, vp], i32),
    "cffm_cfm_attention_dump": ([vp, vp, vp, vp, vp, i32, i32, i32, i32, i32, f32, vp], i32),
    "cffm_cfm_layout": ([vp], i32),
    "cffm_resize_nhwc_to_nchw": ([vp, i32, i64, vp, i32, i32, i32, i32, i32, i32, vp], i32),
    "cffm_resize_argmax": ([vp, vp, i32, i32, i32, i32, i32, i32, vp], i32),
    "cffm_upsample2_argmax": ([vp, i64, vp, i32, i32, i32, i32, i32, i32, i32, i32, vp], i32),
    "cffm_upsample2_argmax_u8": ([vp, i64, vp, i32, i32, i32, i32, i32, i32, i32, i32, vp], i32),
    "cffm_resize_nchw": ([vp, vp, i32, i32, i32, i32, i32, i32, vp], i32),
    "cffm_softmax_nchw": ([vp, vp, i32, i32, i64, vp], i32),
}


class CffmError(RuntimeError):
    pass


def header_symbols(path=HEADER_PATH):
    """Names of every function the header declares."""
    with open(path) as f:
        src = re.sub(r"/\*.*?\*/", "", f.read(), flags=re.S)
    return sorted(set(re.findall(r"\b(cffm_[a-z0-9_]+)\s*\(", src)))


_lock = threading.Lock()
_lib = None
n_launches = 0          # kernels enqueued through this binding (bench.py reports it as gpu_launches)


def load():
    """dlopen the library and bind every symbol; raises CffmError when it is absent."""
    global _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise CffmError(f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                            "(make -C vss_cffm_b200/csrc). There is no fallback implementation.")
        lib = ctypes.CDLL(LIB_PATH)
        for name, (args, res) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.argtypes = args
            fn.restype = res
        if lib.cffm_abi_version() != 1:
            raise CffmError(f"ABI version mismatch: library {lib.cffm_abi_version()}, binding 1")
        _lib = lib
        return lib


def check(status, what):
    if status == 0:
        return
    msg = load().cffm_last_error().decode(errors="replace")
    kind = _STATUS.get(status, f"cudaError {-status}" if status < 0 else f"status {status}")
    raise CffmError(f"{what}: {kind}: {msg}")


launch_hook = None      # optional callable(name, phase, args): phase 0 = before / 1 = after the launch (bench.py)


def call(name, *args):
    """Invoke a compute entry point; counts one kernel launch; raises on a non-zero status."""
    global n_launches
    hook = launch_hook
    if hook is not None:
        hook(name, 0, args)
    st = getattr(load(), name)(*args)
    check(st, name)
    n_launches += 1
    if hook is not None:
        hook(name, 1, args)


def require_device():
    """Raise unless the current CUDA device is an sm_100 part this library can drive."""
    check(load().cffm_device_check(), "cffm_device_check")
