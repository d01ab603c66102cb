"""NVLink peer memory for the row-partitioned SpMM: IPC-shared buffers, a flag barrier and the
halo-pull kernel of libhcspmm (csrc/peer.cu, include/hcspmm.h "multi-GPU").

One process per GPU.  `torch.distributed` is used once per buffer to hand the 64-byte CUDA IPC
handles round (all_gather_object); after that the data path is our own kernels reading the peers'
memory -- no collective per aggregation.
"""
from __future__ import annotations

import ctypes

import torch
import torch.distributed as dist

from . import capi


class _Raw:
    """A device pointer as a __cuda_array_interface__ object (wrapped, not owned, by torch.as_tensor)."""

    def __init__(self, ptr: int, shape, typestr: str):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False),
                                         "version": 2, "strides": None}


class PeerMemory:
    """Buffers every rank of `group` can read, plus a stream-ordered barrier between the ranks."""

    def __init__(self, device: torch.device, group=None):
        self.device, self.group = device, group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self._owned, self._opened = [], []
        self.epoch = 0
        # timeout flag of the barrier kernel in PINNED HOST memory (device-visible at the same address under UVA):
        # the kernel writes it, check() reads it with a plain load -- no synchronisation, so it is tested at every
        # aggregation (hcspmm.dist) and a timed-out barrier stops the job instead of training on stale rows
        self.err = torch.zeros(1, dtype=torch.int32).pin_memory()
        _, self._flag_ptrs = self.shared(max(64, self.world) * 4)
        self.flag_table = torch.tensor(self._flag_ptrs, dtype=torch.int64, device=device)

    def shared(self, nbytes: int):
        """Allocate nbytes here and map every peer's buffer of the same call -> (local ptr, [ptr of rank s]).
        Collective.  A failure on any rank (no IPC, no peer access) raises HcspmmError on EVERY rank, so the
        caller can fall back collectively."""
        L = capi.lib()
        with torch.cuda.device(self.device):
            ptr, handle = ctypes.c_void_p(), ctypes.create_string_buffer(64)
            rc = L.hcspmm_peer_alloc(int(nbytes), ctypes.byref(ptr), handle)
            why = None if rc == 0 else f"hcspmm_peer_alloc: {L.hcspmm_last_error().decode()}"
            if rc == 0:
                self._owned.append(ptr.value)
            ptrs = [ptr.value]
            if self.world > 1:
                handles = [None] * self.world
                dist.all_gather_object(handles, None if why else handle.raw, group=self.group)
                ptrs = []
                for s, h in enumerate(handles):
                    if h is None:
                        why = why or f"rank {s} could not allocate a peer buffer"
                    elif s == self.rank:
                        ptrs.append(ptr.value)
                    elif why is None:
                        p = ctypes.c_void_p()
                        if L.hcspmm_peer_open(ctypes.create_string_buffer(h, 64), ctypes.byref(p)) != 0:
                            why = f"hcspmm_peer_open: {L.hcspmm_last_error().decode()}"
                        else:
                            self._opened.append(p.value)
                            ptrs.append(p.value)
                oks = [None] * self.world
                dist.all_gather_object(oks, why, group=self.group)
                why = next((w for w in oks if w), None)
            if why:
                raise capi.HcspmmError(f"peer memory unavailable: {why}")
        return ptr.value, ptrs

    def tensor(self, ptr: int, shape, dtype=torch.float32) -> torch.Tensor:
        typestr = {torch.float32: "<f4", torch.int32: "<i4", torch.int16: "<i2"}[dtype]
        return torch.as_tensor(_Raw(ptr, shape, typestr), device=self.device)

    def barrier(self):
        """Every rank's work enqueued before its barrier is visible to every rank's work after it."""
        self.epoch += 1
        with torch.cuda.device(self.device):
            capi._check(capi.lib().hcspmm_peer_barrier(self.flag_table.data_ptr(), self.rank, self.world, self.epoch,
                                                       self.err.data_ptr(),
                                                       torch.cuda.current_stream(self.device).cuda_stream),
                        "hcspmm_peer_barrier")

    def check(self):
        v = int(self.err[0])
        if v != 0:
            raise capi.HcspmmError(f"peer barrier timed out waiting for rank {v - 1}: results since then are undefined")

    def close(self):
        L = capi.lib()
        torch.cuda.synchronize(self.device)
        if self.world > 1:
            dist.barrier(group=self.group)       # nobody still reads what is about to be freed
        for p in self._opened:
            L.hcspmm_peer_close(p)
        for p in self._owned:
            L.hcspmm_peer_free(p)
        self._opened, self._owned = [], []


def halo_pull(peer_table: torch.Tensor, lds: int, src_row: torch.Tensor, seg: torch.Tensor, world: int,
              dst: torch.Tensor, col0: int = 0, width: int | None = None, owner_mask: int | None = None,
              first_owner: int = 0, dst_row: torch.Tensor | None = None):
    """hcspmm_halo_pull(_rows) on the current stream: dst[i, col0:col0+width] <- owner's row src_row[i] for the rows of
    the owners in owner_mask (default: all), round-robin from first_owner.  dst_row: list entry i lands in
    dst[dst_row[i]] instead (a subset of the halo: the row-block pipeline)."""
    width = dst.shape[1] - col0 if width is None else width
    owner_mask = (1 << world) - 1 if owner_mask is None else owner_mask
    rows = dst.shape[0] if dst_row is None else int(src_row.numel())
    with torch.cuda.device(dst.device):
        capi._check(capi.lib().hcspmm_halo_pull_rows(peer_table.data_ptr(), lds, src_row.data_ptr(),
                                                     None if dst_row is None else dst_row.data_ptr(), seg.data_ptr(), world,
                                                     owner_mask, first_owner, rows, col0, width, dst.data_ptr(),
                                                     dst.stride(0), torch.cuda.current_stream(dst.device).cuda_stream),
                    "hcspmm_halo_pull_rows")
    return dst


def halo_push(src: torch.Tensor, send_row: torch.Tensor, send_seg: torch.Tensor, dst_table: torch.Tensor, ldd: int,
              world: int, peer_mask: int, first_peer: int = 0, col0: int = 0, width: int | None = None):
    """hcspmm_halo_push on the current stream: peer s's operand rows dst_table[s] + j * ldd <- src[send_row[send_seg[s] + j]]
    for every peer in peer_mask, round-robin from first_peer.  src is FP32 (or a float32 view of bfloat16 rows)."""
    width = src.shape[1] - col0 if width is None else width
    rows = int(send_row.numel())
    with torch.cuda.device(src.device):
        capi._check(capi.lib().hcspmm_halo_push(src.data_ptr(), src.stride(0), send_row.data_ptr(), send_seg.data_ptr(),
                                                dst_table.data_ptr(), ldd, world, peer_mask, first_peer, rows, col0, width,
                                                torch.cuda.current_stream(src.device).cuda_stream),
                    "hcspmm_halo_push")
