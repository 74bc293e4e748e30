"""A raw NCCL communicator through ctypes, for hosts (and tests) that drive `ccst_allreduce_moments` without
torch.distributed.  Plumbing only: the three NCCL calls a C / C++ host would make itself.

    uid = nccl_raw.unique_id()                 # rank 0; ship the 128 bytes to the other ranks by any means
    comm = nccl_raw.comm_init(world, rank, uid)
    _lib.check(_lib.lib().ccst_allreduce_moments(comm, moments.data_ptr(), moments.numel(), stream))
    nccl_raw.comm_destroy(comm)
"""
from __future__ import annotations

import ctypes as C
import glob
import os

_NCCL = None


class _UniqueId(C.Structure):
    _fields_ = [("internal", C.c_ubyte * 128)]  # (a c_char array field would read back truncated at the first NUL)


def _lib():
    global _NCCL
    if _NCCL is None:
        cands = ["libnccl.so.2"]
        try:
            import torch
            cands += glob.glob(os.path.join(os.path.dirname(torch.__file__), "..", "nvidia", "nccl", "lib", "libnccl.so*"))
        except ImportError:
            pass
        err = None
        for c in cands:
            try:
                _NCCL = C.CDLL(c, mode=C.RTLD_GLOBAL)
                break
            except OSError as e:
                err = e
        if _NCCL is None:
            raise RuntimeError(f"NCCL not found ({err})")
        _NCCL.ncclGetErrorString.restype = C.c_char_p
        _NCCL.ncclCommInitRank.argtypes = [C.POINTER(C.c_void_p), C.c_int, _UniqueId, C.c_int]
        _NCCL.ncclCommDestroy.argtypes = [C.c_void_p]
    return _NCCL


def _check(rc: int, what: str):
    if rc != 0:
        raise RuntimeError(f"{what}: {_lib().ncclGetErrorString(rc).decode()}")


def unique_id() -> bytes:
    uid = _UniqueId()
    _check(_lib().ncclGetUniqueId(C.byref(uid)), "ncclGetUniqueId")
    return C.string_at(C.byref(uid), 128)


def comm_init(world: int, rank: int, uid: bytes) -> int:
    """ncclCommInitRank on the CURRENT CUDA device; returns the ncclComm_t as an integer handle."""
    if len(uid) != 128:
        raise ValueError("an ncclUniqueId is 128 bytes")
    u = _UniqueId()
    C.memmove(C.byref(u), uid, 128)
    comm = C.c_void_p()
    _check(_lib().ncclCommInitRank(C.byref(comm), int(world), u, int(rank)), "ncclCommInitRank")
    return comm.value


def comm_destroy(comm: int):
    _check(_lib().ncclCommDestroy(C.c_void_p(comm)), "ncclCommDestroy")
