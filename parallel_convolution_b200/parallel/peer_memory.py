"""Peer-visible device memory for one-process-per-GPU runs.

Each rank allocates its tile buffers, flag pad and status block with plain ``cudaMalloc`` (through
``libpconv_cuda.so``), exports CUDA IPC handles, and all ranks exchange them with one
``all_gather_object`` over the ``torch.distributed`` control plane (NCCL or gloo — it is off the
hot path). A rank then maps only the buffers of its (up to eight) grid neighbours; the resulting
pointers are what the fused kernels dereference — P2P loads/stores over NVLink 5 / NVSwitch.

This is the B200-native replacement for the reference's MPI communicator + derived datatypes
(``mpi/mpi_conv.c:26-28,80-88``): there is no communicator on the data path at all.

``torch.distributed._symmetric_memory`` offers the same capability; plain CUDA IPC is used
because it is a stable public CUDA API, needs no NVSHMEM, and keeps the extension free of any
libtorch dependency. ``PeerGroup.barrier`` is the only host-side synchronisation: once before a
run (every rank's input tile is in place) and once after (safe to overwrite / free).
"""
from __future__ import annotations

from ..ops import cuda as K
from .fused import RankMemory


class PeerGroup:
    """Handle exchange + neighbour mapping for the current process' rank."""

    def __init__(self, rank: int, world: int):
        import torch.distributed as dist
        if not dist.is_initialized():
            raise RuntimeError("torch.distributed must be initialised for multi-process runs")
        self.dist = dist
        self.rank, self.world = rank, world
        self._opened: list[int] = []

    def barrier(self) -> None:
        self.dist.barrier()

    def exchange(self, mem: RankMemory, neighbors: list[int]) -> list[RankMemory | None]:
        """Export ``mem``, gather everybody's handles, map the neighbours. Collective."""
        mine = {"bufs": [K.ipc_export(p) for p in mem.bufs], "ctl": K.ipc_export(mem.flags),
                "arena": K.ipc_export(mem.arena), "nbytes": mem.nbytes}
        table: list[dict | None] = [None] * self.world
        self.dist.all_gather_object(table, mine)
        views: list[RankMemory | None] = []
        cache: dict[int, RankMemory] = {}
        for r in neighbors:
            if r < 0:
                views.append(None)
            elif r == self.rank:
                views.append(mem)
            else:
                if r not in cache:
                    h = table[r]
                    bufs = [K.ipc_open(b) for b in h["bufs"]]
                    ctl = K.ipc_open(h["ctl"])
                    arena = K.ipc_open(h["arena"])
                    self._opened += bufs + [ctl, arena]
                    cache[r] = RankMemory.view(bufs, ctl, h["nbytes"], arena)
                views.append(cache[r])
        return views

    def close(self) -> None:
        for p in self._opened:
            try:
                K.ipc_close(p)
            except Exception:
                pass
        self._opened = []
