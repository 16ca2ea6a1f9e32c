"""CPU ranks with private tiles over the native TCP mesh (``csrc/host/net_ranks.cpp``).

The multi-node form of the CPU backend: what ``HostTile`` does with gloo messages driven from Python (0.5-2 ms per
exchange), done by one native call per ``iterate`` — eight packed strips per exchange over TCP, the part of the
first iteration that reads no halo computed while they are in flight (the reference's schedule,
``mpi/mpi_conv.c:164-248``), k-deep halos. The process group is only used to agree on the rendezvous port.
"""
from __future__ import annotations

import os
import socket

import numpy as np

from ..models.filters import Filter
from ..ops import host as H
from .decomposition import TileGrid

DEFAULT_DEPTH = 6


class NetTile:
    """One rank's private tile. Same surface as ``HostTile`` (load / iterate / result)."""

    def __init__(self, grid: TileGrid, rank: int, filt: Filter, halo_depth: int | None = None, threads: int = 0,
                 dist=None, timeout_ms: float = 120_000.0):
        self.grid, self.rank, self.filt, self.dist = grid, rank, filt, dist
        self.tile = grid.tile(rank)
        self.threads = max(1, int(threads) if threads else 1)
        limit = min(grid.min_tile_rows(), grid.min_tile_cols()) if grid.workers > 1 else 1
        if halo_depth and int(halo_depth) > limit and grid.workers > 1:
            raise ValueError("halo depth exceeds the smallest tile")
        self.k = max(1, min(int(halo_depth) if halo_depth else DEFAULT_DEPTH, limit))
        self.nbrs = grid.neighbors(rank)
        self.exchanges = 0
        t, C = self.tile, grid.channels
        self.pitch = H.net_tile_pitch(t.cols, C, self.k)
        self.bufs = [np.zeros((t.rows + 2 * self.k, self.pitch), dtype=np.uint8) for _ in range(2)]
        self.cur = 0
        self.comm = self._join(float(timeout_ms) / 1e3)

    def _join(self, timeout_s: float) -> H.NetComm:
        world = self.grid.workers
        if self.dist is None or world == 1:
            return H.NetComm(0, 1, "127.0.0.1", 0, -1, timeout_s)
        box = [None]
        listener = None
        if self.rank == 0:                     # rank 0 owns the rendezvous socket: no port race
            listener = socket.socket(socket.AF_INET, socket.SOCK_STREAM)
            listener.setsockopt(socket.SOL_SOCKET, socket.SO_REUSEADDR, 1)
            listener.bind(("", 0))
            listener.listen(world + 8)
            box[0] = listener.getsockname()[1]
        self.dist.broadcast_object_list(box, src=0)
        host = os.environ.get("MASTER_ADDR", "127.0.0.1")
        comm = H.NetComm(self.rank, world, host, int(box[0]), listener.fileno() if listener else -1, timeout_s)
        if listener is not None:
            listener.close()
        return comm

    def _interior(self, b: np.ndarray) -> np.ndarray:
        t, C, k = self.tile, self.grid.channels, self.k
        return b[k:k + t.rows, k * C:(k + t.cols) * C]

    def load(self, tile: np.ndarray) -> None:
        t, C = self.tile, self.grid.channels
        for b in self.bufs:
            b[:] = 0
        self._interior(self.bufs[0])[:] = np.asarray(tile, dtype=np.uint8).reshape(t.rows, t.cols * C)
        self.cur = 0

    def result(self) -> np.ndarray:
        return np.ascontiguousarray(self._interior(self.bufs[self.cur]))

    def iterate(self, loops: int) -> None:
        g, t = self.grid, self.tile
        a, b = self.bufs[self.cur], self.bufs[self.cur ^ 1]
        which, nx = self.comm.run((g.height, g.width, g.channels), (t.row0, t.rows, t.col0, t.cols), self.nbrs, (a, b),
                                  self.filt.spec, loops, self.k, self.threads)
        self.cur ^= which
        self.exchanges += nx

    def close(self) -> None:
        if self.comm is not None:
            self.comm.close()
            self.comm = None
