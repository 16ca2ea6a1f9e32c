"""CPU ranks of one node over a shared image: the host backend without messages.

``HostTile`` (host_backend.py) moves halos as gloo messages, the way the reference moves them as MPI
messages (``mpi/mpi_conv.c:164-248``). When every rank runs on the same node that is copying for its own
sake: ``ShmTile`` keeps both ping-pong images in ONE shared segment (``csrc/host/shm_ranks.cpp``), every rank
computes its tile in place and reads its halo directly from the neighbours' part of the image, and what is
left of the exchange is one epoch counter per rank — the same design the GPU backends use over NVLink peer
memory. ``k`` iterations run per epoch in cache-sized private bands (temporal blocking), OpenMP threads
inside a rank (the reference's ``mpi_omp_conv``).

The process group is only used for the rendezvous (segment name, barriers around loads).
"""
from __future__ import annotations

import os
import socket
import uuid

import numpy as np

from ..models.filters import Filter
from ..ops import host as H
from .decomposition import TileGrid

DEFAULT_DEPTH = 6          # iterations per epoch when the caller does not choose (as pconv_cpu_run_image)


def segment_bytes(grid: TileGrid) -> int:
    """Size of the shared segment for this image: two padded images plus a few pages of counters."""
    C = grid.channels
    pitch = (grid.width * C + 2 * C + 63) // 64 * 64
    return 2 * (((grid.height + 2) * pitch + 4095) // 4096 * 4096) + 4096 * (2 + grid.workers // 64)


def same_node(dist, need_bytes: int = 0) -> bool:
    """True when every rank of the process group runs on this machine and /dev/shm can hold ``need_bytes``
    (shared memory is reachable and large enough; container defaults can be as small as 64 MiB)."""
    try:
        boot = open("/proc/sys/kernel/random/boot_id").read().strip()
    except OSError:
        boot = ""
    room = False
    if os.path.isdir("/dev/shm"):
        try:
            v = os.statvfs("/dev/shm")
            room = v.f_bavail * v.f_frsize >= need_bytes + (16 << 20)
        except OSError:
            room = False
    me = f"{socket.gethostname()}|{boot}"
    names = [None] * dist.get_world_size()
    dist.all_gather_object(names, (me, room))
    return all(n == me and ok for n, ok in names)


class ShmTile:
    """One rank's tile of the shared image. Same surface as ``HostTile`` (load / iterate / result)."""

    def __init__(self, grid: TileGrid, rank: int, filt: Filter, halo_depth: int | None = None,
                 threads: int = 0, dist=None, timeout_ms: float = 60_000.0, segment: H.ShmSegment | None = None):
        self.grid, self.rank, self.filt, self.dist = grid, rank, filt, dist
        self.tile = grid.tile(rank)
        self.threads = max(1, int(threads) if threads else 1)
        limit = min(grid.min_tile_rows(), grid.min_tile_cols()) if grid.workers > 1 else 1 << 30
        self.k = max(1, min(int(halo_depth) if halo_depth else DEFAULT_DEPTH, limit))
        if halo_depth and int(halo_depth) > limit:
            raise ValueError("halo depth exceeds the smallest tile")
        self.timeout_s = float(timeout_ms) / 1e3
        self.nbrs = grid.neighbors(rank)
        self.exchanges = 0
        self.inject_stall = False      # failure-detection tests: this rank stops advancing
        self._own = segment is None
        if segment is not None:
            self.seg = segment
        else:
            self.seg = self._rendezvous()

    def _rendezvous(self) -> H.ShmSegment:
        g, dist = self.grid, self.dist
        if dist is None or g.workers == 1:
            path = f"/dev/shm/pconv_{os.getpid()}_{uuid.uuid4().hex[:12]}"
            seg = H.ShmSegment(path, g.height, g.width, g.channels, g.workers, create=True)
            os.unlink(path)
            return seg
        box = [None]
        seg = None
        if self.rank == 0:
            box[0] = f"/dev/shm/pconv_{os.getpid()}_{uuid.uuid4().hex[:12]}"
            try:
                seg = H.ShmSegment(box[0], g.height, g.width, g.channels, g.workers, create=True)
            except OSError:
                box[0] = None          # tell the others instead of leaving them in the broadcast
        dist.broadcast_object_list(box, src=0)
        if box[0] is None:
            raise OSError("rank 0 could not create the shared image in /dev/shm (no room?); "
                          "use --cpu-transport gloo")
        ok = True
        if self.rank != 0:
            try:
                seg = H.ShmSegment(box[0])
            except OSError:
                ok = False
        oks = [None] * dist.get_world_size()
        dist.all_gather_object(oks, ok)  # also the barrier: every rank has mapped the segment (or failed to)
        if self.rank == 0:
            os.unlink(box[0])          # mapped everywhere: nothing is left behind if a rank dies
        if not all(oks):
            raise OSError("a rank could not attach to the shared image; use --cpu-transport gloo")
        return seg

    # -- tile I/O -----------------------------------------------------------------------------
    def _view(self, which: int) -> np.ndarray:
        g, t, C = self.grid, self.tile, self.grid.channels
        img = self.seg.image(which, g.height, g.row_bytes)
        return img[t.row0:t.row0 + t.rows, t.col0 * C:(t.col0 + t.cols) * C]

    def _barrier(self) -> None:
        if self.dist is not None and self.grid.workers > 1:
            self.dist.barrier()

    def load(self, tile: np.ndarray) -> None:
        t, C = self.tile, self.grid.channels
        self._barrier()                # nobody still reads the old image
        self.seg.reset(self.rank)
        self._view(0)[:] = np.asarray(tile, dtype=np.uint8).reshape(t.rows, t.cols * C)
        self._barrier()                # every tile of image 0 is in place before anyone reads a halo

    def result(self) -> np.ndarray:
        return np.ascontiguousarray(self._view(self.seg.epoch(self.rank)))

    # -- iterations -----------------------------------------------------------------------------
    def iterate(self, loops: int) -> None:
        if self.inject_stall:
            return
        t = self.tile
        self.exchanges += self.seg.run(self.rank, (t.row0, t.row0 + t.rows, t.col0, t.col0 + t.cols), self.nbrs,
                                       self.filt.spec, loops, self.k, self.threads, self.timeout_s)

    def close(self) -> None:
        if self._own and self.seg is not None:
            self.seg.close()
        self.seg = None
