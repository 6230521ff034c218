"""ctypes binding of libvbert_b200.so (the C ABI declared in include/vbert_b200.h).

The library is the product path: there is no Python/PyTorch fallback. If the shared object is
missing or fails to load, importing any compute op raises immediately.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libvbert_b200.so")

VB_EPI_NONE, VB_EPI_GELU, VB_EPI_DGELU = 0, 1, 2

c_void_p, c_int, c_i64, c_f32, c_u64, c_u32 = (
    ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64, ctypes.c_float, ctypes.c_uint64, ctypes.c_uint32)


class GemmArgs(ctypes.Structure):
    """Mirror of vb_gemm_args (include/vbert_b200.h)."""
    _fields_ = [
        ("A", c_void_p), ("lda", c_i64), ("a_mn_major", c_int),
        ("B", c_void_p), ("ldb", c_i64), ("b_mn_major", c_int),
        ("M", c_int), ("N", c_int), ("K", c_int),
        ("D", c_void_p), ("ldd", c_i64),
        ("d_fp32", c_int), ("splits", c_int),
        ("bias", c_void_p),
        ("addend", c_void_p), ("ld_add", c_i64),
        ("epilogue", c_int),
        ("aux_in", c_void_p), ("aux_out", c_void_p), ("ld_aux", c_i64),
        ("dropout_p", c_f32), ("dropout_seed", c_u64), ("dropout_stream", c_u32),
    ]


class VBertLibraryError(RuntimeError):
    pass


_lib = None


def lib():
    """Load (once) and return the ctypes handle; raise loudly when the CUDA library is absent."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise VBertLibraryError(
                f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'`. "
                "visualbert_b200 has no CPU or PyTorch fallback.")
        h = ctypes.CDLL(LIB_PATH)
        h.vb_last_error.restype = ctypes.c_char_p
        h.vb_launch_count.restype = ctypes.c_int64
        h.vb_abi_version.restype = ctypes.c_int
        _lib = h
    return _lib


def check(rc, what):
    if rc != 0:
        msg = lib().vb_last_error().decode("utf-8", "replace")
        raise VBertLibraryError(f"{what} failed (status {rc}): {msg}")


def launch_count():
    return int(lib().vb_launch_count())
