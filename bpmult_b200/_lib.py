"""ctypes binding of libbpmult_b200.so (the C ABI declared in include/bpmult_b200.h).

The product path has no CPU fallback: if the CUDA library is missing, or a call is made without an sm_100 GPU,
this module raises.  `python -m bpmult_b200.build` (or `__graft_entry__.build()`) compiles the library in-tree."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libbpmult_b200.so")

BPM_F32, BPM_BF16 = 0, 1


class Dropout(C.Structure):
    _fields_ = [("seed", C.c_uint64), ("seed_ptr", C.c_void_p), ("site", C.c_uint64), ("p", C.c_float)]


class Gemm(C.Structure):
    _fields_ = [("ab_dtype", C.c_int), ("ta", C.c_int), ("tb", C.c_int), ("M", C.c_int), ("N", C.c_int), ("K", C.c_int),
                ("A", C.c_void_p), ("lda", C.c_int), ("B", C.c_void_p), ("ldb", C.c_int),
                ("C", C.c_void_p), ("ldc", C.c_int), ("c_dtype", C.c_int),
                ("bias", C.c_void_p), ("alpha", C.c_float), ("act", C.c_int),
                ("drop", Dropout),
                ("gate", C.c_void_p), ("ldg", C.c_int), ("gate_dtype", C.c_int), ("gate_scale", C.c_float),
                ("residual", C.c_void_p), ("ldr", C.c_int), ("res_dtype", C.c_int),
                ("accumulate", C.c_int), ("split_k", C.c_int), ("colsum_out", C.c_void_p)]


class Attn(C.Structure):
    _fields_ = [("dtype", C.c_int), ("B", C.c_int), ("T", C.c_int), ("S", C.c_int), ("H", C.c_int), ("dh", C.c_int),
                ("dhp", C.c_int), ("mask_off", C.c_int), ("key_pad", C.c_void_p), ("drop", Dropout), ("drop_bits", C.c_void_p),
                ("ld_kv", C.c_int), ("ld_dkv", C.c_int)]


_P, _I, _F, _L = C.c_void_p, C.c_int, C.c_float, C.c_int64
# name -> argtypes  (every function returns int unless listed in _RESTYPE)
SIGNATURES = {
    "bpm_version": [],
    "bpm_last_error": [],
    "bpm_device_ok": [_I],
    "bpm_debug_set": [_I, _I],
    "bpm_debug_set_ptr": [_P],
    "bpm_pack_matrix": [_P, _I, _I, _I, _P, _I, _I, _I, _I, _I, _I, _I, _P],
    "bpm_unpack_matrix": [_P, _I, _I, _P, _I, _I, _I, _I, _I, _I, _I, _I, _F, _P],
    "bpm_remap_batch": [_P, _I, _I, _P],
    "bpm_remap_units": [_P, _P, _I, _I, _P],
    "bpm_stage_rows": [_P, _I, _I, _I, _L, _L, _L, _P, _I, _I, _I, Dropout, _P],
    "bpm_unstage_rows": [_P, _I, _I, _I, _I, _I, _P, _L, _L, _L, _I, Dropout, _P],
    "bpm_embed_fwd": [_P, _I, _P, _I, _I, _I, _I, _F, _P, _I, Dropout, _P],
    "bpm_embed_bwd": [_P, _I, _I, _I, _F, _P, _I, Dropout, _P],
    "bpm_layernorm_fwd": [_P, _I, _P, _P, _I, _I, _I, _F, _P, _I, _P, _P, _P],
    "bpm_layernorm_bwd": [_P, _I, _P, _I, _P, _P, _P, _I, _I, _I, _P, _I, _P, _P, _P],
    "bpm_layernorm_bwd_cast": [_P, _I, _P, _I, _P, _P, _P, _I, _I, _I, _P, _I, _P, _P, _P, _I, Dropout, _P],
    "bpm_ln_fold_fwd": [_P, _I, _P, _P, _P, _I, _I, _I, _I, _P, _I, _I, _P, _P],
    "bpm_ln_fold_bwd": [_P, _I, _P, _P, _I, _I, _I, _I, _P, _I, _P, _P, _I, _P, _P, _P, _P],
    "bpm_ln_fold_batch": [_P, _I, _I, _I, _P],
    "bpm_gemm": [C.POINTER(Gemm), _P],
    "bpm_colsum": [_P, _I, _I, _I, _I, _P, _P],
    "bpm_xattn_fwd": [C.POINTER(Attn), _P, _P, _P, _P, _P, _P],
    "bpm_xattn_bwd": [C.POINTER(Attn), _P, _P, _P, _P, _P, _P, _P, _P, _F, _P, _P, _P],
    "bpm_xattn_bwd_workspace": [C.POINTER(Attn)],
    "bpm_xattn_weights": [C.POINTER(Attn), _P, _P, _P, _P, _P],
    "bpm_gmu_fwd": [_I, _I, _P, _P, _P, _P, _P, _P, _I, _I, _P, _P, _P],
    "bpm_gmu_bwd": [_I, _I, _P, _P, _P, _P, _P, _P, _I, _I, _P, _P, _P, _P, _P, _P],
    "bpm_add": [_I, _P, _P, _P, _L, _P],
    "bpm_axpy_f32": [_P, _I, _P, _L, _I, _P],
    "bpm_cast_drop": [_P, _P, _I, _I, _I, Dropout, _P],
    "bpm_pool_fwd": [_P, _I, _I, _I, _I, _P, _I, _I, _P],
    "bpm_pool_bwd": [_P, _I, _I, _I, _I, _I, _P, _P],
    "bpm_tsgate_fwd": [_P, _P, _I, _I, _I, _P, _P, _P],
    "bpm_tsgate_bwd": [_P, _P, _P, _I, _I, _I, _P, _P, _P],
    "bpm_bce_fwd_bwd": [_P, _I, _P, _P, _I, _I, _F, _P, _P, _P],
    "bpm_timelin_fwd": [_I, _P, _P, _P, _P, _I, _I, _I, _I, _I, _P],
    "bpm_timelin_bwd": [_I, _P, _P, _P, _P, _I, _P, _P, _I, _I, _I, _I, _I, _P],
    "bpm_conv1d_im2col": [_P, _I, _I, _I, _I, _I, _I, _I, _P, _I, _P],
    "bpm_conv1d_col2im": [_P, _I, _I, _I, _I, _I, _I, _I, _P, _I, _P],
    "bpm_conv1d_pack_weight": [_P, _I, _I, _I, _P, _I, _P],
    "bpm_conv1d_unpack_wgrad": [_P, _I, _I, _I, _P, _I, _P],
    "bpm_adaptive_pool_fwd": [_P, _I, _I, _I, _I, _I, _I, _P, _I, _I, _P],
    "bpm_adaptive_pool_bwd": [_P, _I, _I, _I, _I, _I, _P, _I, _P],
    "bpm_adam_step": [_P, _P, _P, _P, _L, _F, _F, _F, _F, _F, _P, _P, _P],
}
_RESTYPE = {"bpm_last_error": C.c_char_p, "bpm_xattn_bwd_workspace": C.c_int64}

_lib = None


class BpmError(RuntimeError):
    pass


def load():
    """Loads the CUDA library; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise BpmError("bpmult_b200: %s not found -- build it with `python -m bpmult_b200.build` "
                       "(nvcc, sm_100a).  There is no CPU fallback." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, args in SIGNATURES.items():
        fn = getattr(lib, name)             # AttributeError if the library does not export a declared symbol
        fn.argtypes = args
        fn.restype = _RESTYPE.get(name, C.c_int)
    _lib = lib
    return lib


def check(rc, what=""):
    if rc != 0:
        msg = load().bpm_last_error()
        raise BpmError("bpmult_b200 %s failed (rc=%d): %s" % (what, rc, msg.decode() if msg else "?"))
