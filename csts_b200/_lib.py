"""ctypes binding of libcsts_b200.so — the C-ABI declared in include/csts_b200.h.

PyTorch is used for device memory and streams only: every wrapper passes raw device pointers,
sizes and the current CUDA stream.  There is no fallback: if the shared library is missing or a
kernel reports an error, a RuntimeError is raised.
"""
import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
# CSTS_B200_LIB selects another build of the same library (A/B tuning runs only)
LIB_PATH = os.environ.get("CSTS_B200_LIB") or os.path.join(_HERE, "libcsts_b200.so")

F32, BF16, F16 = 0, 1, 2
_DT = {torch.float32: F32, torch.bfloat16: BF16, torch.float16: F16}


class GemmArgs(C.Structure):
    _fields_ = [
        ("A", C.c_void_p), ("B", C.c_void_p), ("C", C.c_void_p), ("Z", C.c_void_p),
        ("bias", C.c_void_p), ("residual", C.c_void_p), ("row_scale", C.c_void_p), ("rowsum", C.c_void_p),
        ("lda", C.c_int64), ("ldb", C.c_int64), ("ldc", C.c_int64), ("ldz", C.c_int64), ("ldr", C.c_int64),
        ("sA1", C.c_int64), ("sA2", C.c_int64), ("sB1", C.c_int64), ("sB2", C.c_int64), ("sC1", C.c_int64), ("sC2", C.c_int64),
        ("M", C.c_int32), ("N", C.c_int32), ("K", C.c_int32),
        ("batch1", C.c_int32), ("batch2", C.c_int32),
        ("a_kmajor", C.c_int32), ("b_kmajor", C.c_int32),
        ("c_dtype", C.c_int32), ("act", C.c_int32), ("accumulate", C.c_int32), ("res_mod", C.c_int32),
        ("split_k", C.c_int32), ("alpha", C.c_float), ("backend", C.c_int32), ("rows_per_scale", C.c_int32),
        ("a_dtype", C.c_int32), ("b_dtype", C.c_int32), ("z_dtype", C.c_int32),
        ("tile_n", C.c_int32), ("ctas", C.c_int32),
        ("rowvec", C.c_void_p),
    ]


class PoolArgs(C.Structure):
    _fields_ = [
        ("inp", C.c_void_p), ("out", C.c_void_p), ("w", C.c_void_p), ("gamma", C.c_void_p), ("beta", C.c_void_p),
        ("pre", C.c_void_p), ("mean", C.c_void_p), ("rstd", C.c_void_p),
        ("in_sB", C.c_int64), ("in_sH", C.c_int64), ("in_sP", C.c_int64),
        ("out_sB", C.c_int64), ("out_sH", C.c_int64), ("out_sP", C.c_int64),
        ("B", C.c_int32), ("heads", C.c_int32), ("d", C.c_int32),
        ("Ti", C.c_int32), ("Hi", C.c_int32), ("Wi", C.c_int32),
        ("To", C.c_int32), ("Ho", C.c_int32), ("Wo", C.c_int32),
        ("st", C.c_int32), ("sh", C.c_int32), ("sw", C.c_int32),
        ("transposed", C.c_int32), ("eps", C.c_float), ("dtype", C.c_int32),
        ("in2", C.c_void_p), ("out2", C.c_void_p), ("w2", C.c_void_p), ("gamma2", C.c_void_p), ("beta2", C.c_void_p),
        ("pre2", C.c_void_p), ("mean2", C.c_void_p), ("rstd2", C.c_void_p),
    ]


class WgradArgs(C.Structure):
    _fields_ = [
        ("small", C.c_void_p), ("big", C.c_void_p), ("dw", C.c_void_p),
        ("small_sB", C.c_int64), ("small_sH", C.c_int64), ("small_sP", C.c_int64),
        ("big_sB", C.c_int64), ("big_sH", C.c_int64), ("big_sP", C.c_int64),
        ("B", C.c_int32), ("heads", C.c_int32), ("d", C.c_int32),
        ("Ts", C.c_int32), ("Hs", C.c_int32), ("Ws", C.c_int32),
        ("Tb", C.c_int32), ("Hb", C.c_int32), ("Wb", C.c_int32),
        ("st", C.c_int32), ("sh", C.c_int32), ("sw", C.c_int32),
        ("small_dtype", C.c_int32), ("big_dtype", C.c_int32),
        ("small2", C.c_void_p), ("big2", C.c_void_p), ("dw2", C.c_void_p),
    ]


# name -> argtypes (every function returns int; the trailing void* is the cudaStream_t)
_P, _I, _L, _F = C.c_void_p, C.c_int, C.c_int64, C.c_float
SIGNATURES = {
    "csts_version": [],
    "csts_last_error": [C.c_char_p, _I],
    "csts_check_device": [],
    "csts_gemm": [C.POINTER(GemmArgs), _P],
    "csts_layernorm_fwd": [_P, _I, _P, _I, _P, _P, _P, _P, _L, _I, _F, _P],
    "csts_layernorm_bwd": [_P, _I, _P, _I, _P, _P, _P, _P, _P, _I, _P, _P, _L, _I, _P, _I, _P, _I, _P],
    "csts_layernorm_bwd_pair": [_P, _P, _I, _P, _P, _I, _P, _P, _P, _P, _P, _P, _P, _P, _I, _P, _P, _P, _P, _L, _I, _P],
    "csts_rowdot": [_P, _P, _I, _P, _I, _I, _I, _I, _P],
    "csts_softmax_fwd": [_P, _P, _I, _L, _I, _I, _I, _I, _I, _I, _P],
    "csts_softmax_bwd": [_P, _I, _P, _P, _I, _L, _I, _I, _I, _F, _P],
    "csts_cast16": [_P, _P, _I, _L, _I, _I, _P, _I, _P],
    "csts_permute_021": [_P, _P, _I, _I, _I, _I, _P],
    "csts_add_f32": [_P, _P, _P, _L, _P],
    "csts_scale_f32": [_P, _P, _P, _L, _P],
    "csts_colsum": [_P, _I, _P, _L, _I, _L, _P],
    "csts_dwconv": [C.POINTER(PoolArgs), _P],
    "csts_dwconv_wgrad": [C.POINTER(WgradArgs), _P],
    "csts_maxpool_fwd": [_P, _P, _P, _I, _I, _I, _I, _I, _P],
    "csts_maxpool_bwd": [_P, _P, _P, _I, _I, _I, _I, _I, _P],
    "csts_upsample_fwd": [_P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _P],
    "csts_upsample_bwd": [_P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _I, _P],
    "csts_im2col_patch": [_P, _P, _I, _I, _I, _I, _I, _I, _I, _P],
    "csts_pos_embed": [_P, _P, _P, _I, _I, _I, _P],
    "csts_pos_embed_bwd": [_P, _P, _P, _I, _I, _I, _I, _P],
    "csts_reweight_fwd": [_P, _P, _P, _I, _I, _I, _I, _L, _P],
    "csts_reweight_bwd": [_P, _P, _P, _P, _P, _I, _I, _I, _I, _L, _P],
    "csts_token_mean_fwd": [_P, _P, _I, _I, _I, _I, _P],
    "csts_token_mean_bwd": [_P, _P, _I, _I, _I, _I, _P],
    "csts_classifier_fwd": [_P, _P, _P, _P, _P, _I, _I, _I, _I, _P],
    "csts_classifier_bwd": [_P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _P],
    "csts_kldiv_frame_softmax": [_P, _P, _P, _P, _P, _P, _I, _I, _I, _F, _P],
    "csts_sim_matrix_fwd": [_P, _P, _P, _P, _P, _I, _I, _F, _P],
    "csts_sim_matrix_bwd": [_P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _P],
    "csts_egonce": [_P, _P, _P, _P, _I, _F, _P],
    "csts_adaptive_f1": [_P, _P, _P, _P, _I, _I, _I, _I, _I, _P, _P, _P],
    "csts_grad_sqnorm": [_P, _P, _I, _P, _P],
    "csts_clip_adamw_step": [_P, _P, _I, _P, _P, _P, _P, _P, _P, C.c_double, C.c_double, _F, _F, _P],
}

_lib = None


def load():
    """dlopen the library and bind every symbol of the C-ABI (fails loudly if one is missing)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(make -C csts_b200/csrc).  csts_b200 has no CPU or PyTorch fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, argtypes in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the .so does not export it
        fn.argtypes = argtypes
        fn.restype = C.c_int
    lib.csts_gemm_backend.argtypes = [C.POINTER(GemmArgs)]
    lib.csts_gemm_backend.restype = C.c_int
    lib.csts_gemm_plan.argtypes = [C.POINTER(GemmArgs), C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]
    lib.csts_gemm_plan.restype = C.c_int
    lib.csts_mt_chunk_elems.argtypes = []
    lib.csts_mt_chunk_elems.restype = C.c_int
    lib.csts_launch_count.argtypes = [C.c_int]
    lib.csts_launch_count.restype = C.c_longlong
    _lib = lib
    return lib


def launch_count(reset=False) -> int:
    """Kernels launched by libcsts_b200 since load / last reset."""
    return int(load().csts_launch_count(1 if reset else 0))


def last_error() -> str:
    buf = C.create_string_buffer(512)
    load().csts_last_error(buf, 512)
    return buf.value.decode(errors="replace")


def check(rc: int, what: str):
    if rc != 0:
        raise RuntimeError(f"{what} failed (code {rc}): {last_error()}")


def stream_ptr():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def dt(t):
    return _DT[t.dtype]


_device_checked = False
PROFILE = None        # when a list: every call() appends (name, start_event, end_event) — tools/step_profile.py


def _describe(args):
    """Shape signature of a call for tools/step_profile.py: the plain integer arguments, or the geometry fields of an argument block."""
    out = [a for a in args if isinstance(a, int) and not isinstance(a, bool)]
    for a in args:
        s = getattr(a, "_obj", None)
        if isinstance(s, PoolArgs):
            out += [s.B, s.heads, s.d, s.Ti, s.Hi, s.Wi, s.To, s.Ho, s.Wo, s.transposed, int(bool(s.gamma)), int(bool(s.in2))]
        elif isinstance(s, WgradArgs):
            out += [s.B, s.heads, s.d, s.Ts, s.Hs, s.Ws, s.Tb, s.Hb, s.Wb, int(bool(s.small2))]
        elif isinstance(s, GemmArgs):
            out += [s.M, s.N, s.K, s.batch1 * s.batch2, s.a_kmajor, s.b_kmajor, s.act, s.c_dtype, int(bool(s.residual)), s.split_k]
    return str(tuple(out))


def call(name, *args):
    """Invoke `name(*args, stream)` on the current CUDA stream and raise on a non-zero code."""
    global _device_checked
    lib = load()
    if not _device_checked:
        if not torch.cuda.is_available():
            raise RuntimeError("csts_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        check(lib.csts_check_device(), "csts_check_device")
        _device_checked = True
    if PROFILE is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        check(getattr(lib, name)(*args, stream_ptr()), name)
        e1.record()
        PROFILE.append((name + _describe(args), e0, e1))
        return
    check(getattr(lib, name)(*args, stream_ptr()), name)
