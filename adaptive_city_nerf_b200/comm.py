"""ctypes front-end of libacn_b200_comm.so (include/acn_b200_comm.h): the NCCL-backed exchange entries a non-PyTorch
host uses (acn_comm_init / acn_allreduce / acn_allgather / acn_alltoall_samples).  The package's own multi-GPU paths
(distributed.py) go through torch.distributed and symmetric memory; this module exists so the C ABI is exercised by the
tests the same way a C host would call it -- raw device pointers, host count arrays, a CUDA stream."""
from __future__ import annotations

import ctypes as C
from pathlib import Path
from typing import Sequence

import torch

from . import _lib

LIB_PATH = Path(__file__).resolve().parent / "libacn_b200_comm.so"
HEADER = _lib.HEADER.parent / "acn_b200_comm.h"
ID_BYTES = 128
OP_SUM, OP_MAX = 0, 1

_clib = None


def comm_lib():
    global _clib
    if _clib is None:
        if not LIB_PATH.exists():
            raise RuntimeError(f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'`")
        l = C.CDLL(str(LIB_PATH))
        for name, argtypes in _lib._parse_header(HEADER).items():
            fn = getattr(l, name)
            fn.argtypes = argtypes
            fn.restype = C.c_char_p if name == "acn_comm_last_error" else C.c_int
        _clib = l
    return _clib


def _check(rc: int) -> None:
    if rc != 0:
        msg = comm_lib().acn_comm_last_error()
        raise RuntimeError(f"libacn_b200_comm error {rc}: {msg.decode() if msg else '?'}")


def unique_id() -> bytes:
    """Rank 0: the 128-byte NCCL rendezvous id, to be shipped to the other ranks over any host channel."""
    buf = C.create_string_buffer(ID_BYTES)
    _check(comm_lib().acn_comm_unique_id(buf))
    return buf.raw


class Communicator:
    def __init__(self, device: torch.device, uid: bytes, rank: int, world: int):
        assert len(uid) == ID_BYTES
        self.device = torch.device(device)
        self.rank, self.world = rank, world
        h = C.c_void_p()
        _check(comm_lib().acn_comm_init(self.device.index, uid, rank, world, C.byref(h)))
        self._h = h

    def close(self) -> None:
        if self._h:
            comm_lib().acn_comm_destroy(self._h)
            self._h = None

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def allreduce_(self, t: torch.Tensor, op: int = OP_SUM) -> torch.Tensor:
        assert t.is_cuda and t.is_contiguous() and t.dtype in (torch.float32, torch.float16)
        _check(comm_lib().acn_allreduce(self._h, C.c_void_p(t.data_ptr()), t.numel(), _lib.F32 if t.dtype == torch.float32 else _lib.F16,
                                        op, self._stream()))
        return t

    def allgather(self, t: torch.Tensor) -> torch.Tensor:
        assert t.is_cuda and t.is_contiguous()
        out = torch.empty((self.world,) + tuple(t.shape), dtype=t.dtype, device=t.device)
        _check(comm_lib().acn_allgather(self._h, C.c_void_p(t.data_ptr()), C.c_void_p(out.data_ptr()), t.numel() * t.element_size(),
                                        self._stream()))
        return out

    def alltoall_samples(self, send: torch.Tensor, send_counts: Sequence[int], recv_counts: Sequence[int]) -> torch.Tensor:
        """send (sum(send_counts), row...) rows grouped by destination rank -> (sum(recv_counts), row...) grouped by source."""
        assert send.is_cuda and send.is_contiguous() and len(send_counts) == self.world == len(recv_counts)
        row_bytes = (send[0].numel() if send.shape[0] else int(torch.tensor(send.shape[1:]).prod())) * send.element_size()
        recv = torch.empty((int(sum(recv_counts)),) + tuple(send.shape[1:]), dtype=send.dtype, device=send.device)
        sc = (C.c_int64 * self.world)(*[int(c) for c in send_counts])
        rcv = (C.c_int64 * self.world)(*[int(c) for c in recv_counts])
        _check(comm_lib().acn_alltoall_samples(self._h, C.c_void_p(send.data_ptr()), sc, C.c_void_p(recv.data_ptr()), rcv, row_bytes,
                                               self._stream()))
        return recv
