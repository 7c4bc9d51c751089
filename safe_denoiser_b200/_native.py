"""ctypes binding of libsdn_repel.so (the C ABI declared in include/sdn_repel.h).

There is no CPU fallback: if the library is missing, or a tensor is not on a
CUDA device, the calls raise.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SDN_REPEL_LIB") or os.path.join(_HERE, "libsdn_repel.so")   # the override is for A/B experiments

PATH_AUTO, PATH_GENERIC, PATH_STREAM, PATH_UMMA, PATH_UMMA_BF16, PATH_FLASH = 0, 1, 2, 3, 4, 5
EPI_GATE, EPI_RETURN_NEG = 1, 2

_lib = None

_f, _i32, _i64, _p, _sz = C.c_float, C.c_int32, C.c_int64, C.c_void_p, C.c_size_t

# name -> (restype, argtypes); mirrors include/sdn_repel.h one to one
SIGNATURES = {
    "sdn_abi_version": (C.c_int, []),
    "sdn_error_string": (C.c_char_p, [C.c_int]),
    "sdn_launch_count": (C.c_uint64, []),
    "sdn_set_option": (C.c_int, [_i32, _i32]),
    "sdn_debug_read": (_i32, [C.POINTER(C.c_uint32), _i32]),
    "sdn_debug_trace_read": (_sz, [_p, _sz]),
    "sdn_debug_accum_trace_read": (_sz, [_p, _sz]),
    "sdn_profile_enable": (None, [_i32]),
    "sdn_profile_read": (_i32, [_i32, C.c_char_p, _i32, C.POINTER(C.c_float)]),
    "sdn_bank_prepare": (C.c_int, [_p, _i64, _i64, _p, _p, _p]),
    "sdn_bank_build": (C.c_int, [_p, _i64, _i32, _i64, _p, _p, _p, _p]),
    "sdn_query_prepare": (C.c_int, [_p, _p, _f, _f, _i64, _i64, _i32, _p, _p, _p, _p]),
    "sdn_repel_workspace_bytes": (_sz, [_i64, _i64, _i64, _i32]),
    "sdn_repel_path": (_i32, [_i64, _i64, _i64, _i32, _i32]),
    "sdn_repel_partial": (C.c_int, [_p, _p, _p, _i64, _i64, _p, _p, _i64, _f, _i32, _f,
                                    _p, _p, _p, _p, _sz, _i32, _p]),
    "sdn_epilogue_correct": (C.c_int, [_p, _p, _i64, _i64, _f, _f, _f, _i32, _p, _p, _p, _p, _p, _p]),
    "sdn_epilogue_ddpm": (C.c_int, [_p, _p, _i64, _i64, _f, _f, _f, _i32, _p, _p, _p, _p,
                                    _f, _f, _f, _f, _f, _p, _p, _p, _p, _p, _p]),
    "sdn_epilogue_ddim": (C.c_int, [_p, _p, _i64, _i64, _f, _f, _f, _i32, _p, _p, _p,
                                    _f, _f, _f, _f, _p, _p, _p, _p, _p, _p]),
    "sdn_epilogue_flow": (C.c_int, [_p, _p, _i64, _i64, _f, _f, _p, _p, _p, _f, _f, _p, _p, _p, _p, _p]),
    "sdn_conditioning_fused": (C.c_int, [_p, _p, _p, _i64, _i64, _p, _i64, _f, _i32, _f, _f, _f, _f, _i32,
                                         _p, _p, _p, _p, _p, _p, _p, _p, _sz, _i32, _p]),
    "sdn_shard_merge_correct": (C.c_int, [_p, _p, _p, _i32, _i32, _p, _i64, _i64, _f, _f, _f, _i32,
                                          _p, _p, _p, _p, _p]),
    "sdn_sparse_repel": (C.c_int, [_p, _p, _i64, _i64, _p, _p, _p, _i64, _f, _f, _p, _p, _p, _sz, _p]),
    "sdn_sparse_partial": (C.c_int, [_p, _p, _i64, _i64, _p, _p, _i64, _f, _p, _p, _p, _sz, _p]),
    "sdn_sparse_apply": (C.c_int, [_p, _p, _i64, _i64, _f, _p, _p, _p, _p]),
    "sdn_sparse_partial_planes": (C.c_int, [_p, _p, _i64, _i64, _p, _p, _i64, _f, _p, _p, _p, _sz, _p]),
    "sdn_conditioning_host": (C.c_int, [_p, _p, _p, _i64, _i64, _p, _i64, _i32, _f, _i32, _f, _f, _f,
                                        _p, _i32, _p]),
    "sdn_host_release": (None, []),
    "sdn_host_pipe_create": (C.c_int, [_i64, _i64, _i64, _i32, _p]),
    "sdn_host_pipe_submit": (C.c_int, [_p, _i32, _p, _p, _p, _p, _p, _p, _f, _i32, _f, _f, _f]),
    "sdn_host_pipe_wait": (C.c_int, [_p, _i32]),
    "sdn_host_pipe_destroy": (None, [_p]),
}


class NativeLibraryMissing(RuntimeError):
    pass


def lib():
    """Load (once) and return the C-ABI library.  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise NativeLibraryMissing(
            f"{LIB_PATH} not found: build it with `python -m safe_denoiser_b200.build` "
            "(there is no CPU or PyTorch fallback for the repellency projection)")
    handle = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(handle, name)   # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    if handle.sdn_abi_version() != 1:
        raise RuntimeError("libsdn_repel.so ABI version mismatch")
    _lib = handle
    return _lib


def check(rc):
    if rc != 0:
        msg = lib().sdn_error_string(rc).decode()
        raise RuntimeError(f"sdn_repel call failed ({rc}): {msg}")


def debug_read():
    """Diagnostic record of a one-pass kernel that trapped on a bounded wait: [code, cta, thread, tile, extra] or None."""
    buf = (C.c_uint32 * 8)()
    if not lib().sdn_debug_read(buf, 8):
        return None
    return [int(v) for v in buf]


def ptr(t):
    """Device pointer of a CUDA tensor (None -> NULL)."""
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError("repellency kernels need CUDA tensors; there is no CPU fallback")
    if not t.is_contiguous():
        raise RuntimeError("repellency kernels need contiguous tensors")
    return t.data_ptr()


def current_stream():
    import torch
    return torch.cuda.current_stream().cuda_stream


OPT_SKIP_NEGLIGIBLE = 1


def set_option(key, value):
    check(lib().sdn_set_option(int(key), int(value)))


def profile_enable(on=True):
    lib().sdn_profile_enable(1 if on else 0)


def profile_read():
    """[(kernel name, ms)] of the last sdn_repel_partial call (needs profile_enable(True) before it)."""
    out, i = [], 0
    buf = C.create_string_buffer(64)
    ms = C.c_float(0.0)
    while lib().sdn_profile_read(i, buf, 64, C.byref(ms)):
        out.append((buf.value.decode(), float(ms.value)))
        i += 1
    return out


def launch_count():
    return int(lib().sdn_launch_count())
