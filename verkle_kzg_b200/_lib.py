"""ctypes binding of libvkzg.so (the C ABI in include/vkzg.h).

The library is the product: there is no Python or CPU implementation behind it.  Importing this module
on a box where the shared object is missing, or calling into it without an sm_100 GPU, fails loudly.
"""
import ctypes
import os
import subprocess

import numpy as np

_PKG = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_PKG, "libvkzg.so")
_CSRC = os.path.join(_PKG, "csrc")

OK = 0
KEY_WINDOW = 1
KEY_MSM = 2


class VkzgError(RuntimeError):
    def __init__(self, status, what):
        self.status = status
        super().__init__(f"{what}: status {status} ({strerror(status)})")


def build(force=False, jobs=8):
    """Compile libvkzg.so for sm_100a in-tree (nvcc cross-compiles without a GPU)."""
    if force:
        subprocess.check_call(["make", "-C", _CSRC, "clean"], stdout=subprocess.DEVNULL)
    subprocess.check_call(["make", "-C", _CSRC, f"-j{jobs}"], stdout=subprocess.DEVNULL)
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            raise ImportError(
                f"{_SO} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(verkle_kzg_b200 has no CPU fallback)")
        _lib = ctypes.CDLL(_SO)
        _lib.vkzg_strerror.restype = ctypes.c_char_p
        _lib.vkzg_ctx_launches.restype = ctypes.c_uint64
        _lib.vkzg_key_table_bytes.restype = ctypes.c_uint64
        _lib.vkzg_abi_version.restype = ctypes.c_uint32
    return _lib


def strerror(status):
    return lib().vkzg_strerror(ctypes.c_int32(status)).decode()


def check(status, what):
    if status != OK:
        raise VkzgError(status, what)


def hptr(a):
    """host pointer of a C-contiguous numpy array (None -> NULL)"""
    if a is None:
        return None
    assert isinstance(a, np.ndarray) and a.flags["C_CONTIGUOUS"], "need a C-contiguous numpy array"
    return a.ctypes.data_as(ctypes.c_void_p)


def dptr(t):
    """device pointer of a torch CUDA tensor / int (None -> NULL)"""
    if t is None:
        return None
    if isinstance(t, int):
        return ctypes.c_void_p(t)
    assert t.is_cuda and t.is_contiguous()
    return ctypes.c_void_p(t.data_ptr())


def u8(a, last):
    a = np.ascontiguousarray(a, dtype=np.uint8)
    assert a.shape[-1] == last, f"expected trailing dimension {last}, got {a.shape}"
    return a
