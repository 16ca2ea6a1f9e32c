"""Host (CPU) backend: native C++ compute + ``torch.distributed`` (gloo) halo exchange.

This is the framework's analogue of the reference's ``mpi_conv`` / ``mpi_omp_conv`` programs
(``mpi/mpi_conv.c:164-248``): one process per tile, padded ping-pong tile buffers, non-blocking
neighbour exchange overlapped with the interior compute, OpenMP threads inside a rank
(``--threads``; the reference's ``OMP_NUM_THREADS``). Differences by design:

* all **eight** neighbours are exchanged, so np>1 output equals np=1 output (the reference skips
  the diagonals and never recomputes some corner pixels: defects D1/D2);
* the halo can be ``k`` pixels deep with one exchange per ``k`` iterations (temporal blocking);
  ``k=1`` reproduces the reference's schedule: post → interior → wait → border ring;
* remainder tiles are supported.

It exists for plumbing (BASELINE config 1 runs with no GPU), as a semantic cross-check of the GPU
backends, and for CPU-only boxes. World size 1 needs no process group at all.
"""
from __future__ import annotations

import os

import numpy as np

from ..models.filters import Filter
from ..ops import host as H
from .decomposition import NEIGHBOR_OFFSETS, TileGrid


class HostTile:
    """One rank's padded tile pair and its halo exchange."""

    def __init__(self, grid: TileGrid, rank: int, filt: Filter, halo_depth: int = 1,
                 threads: int = 0, dist=None):
        self.grid, self.rank, self.filt = grid, rank, filt
        self.k = max(1, int(halo_depth))
        self.threads, self.dist = threads, dist
        self.tile = grid.tile(rank)
        t, C = self.tile, grid.channels
        if grid.workers > 1 and self.k > min(grid.min_tile_rows(), grid.min_tile_cols()):
            raise ValueError("halo depth exceeds the smallest tile")
        self.pitch = ((t.cols + 2 * self.k) * C + 63) // 64 * 64
        self.bufs = [np.zeros((t.rows + 2 * self.k, self.pitch), dtype=np.uint8) for _ in range(2)]
        self.cur = 0
        self.nbrs = grid.neighbors(rank)
        self.exchanges = 0
        # Debug aid: overwrite every halo that has a neighbour with 0xA5 right before
        # each exchange, so a strip that is consumed without having been freshly received (a missing
        # diagonal, a wrong tag, a skipped wait: the reference's defects D1/D2) corrupts the output
        # instead of silently reusing last iteration's values.
        self.poison = os.environ.get("PCONV_POISON_HALO", "0") not in ("", "0")

    # -- tile I/O -----------------------------------------------------------------------------
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

    # -- halo exchange --------------------------------------------------------------------------
    def _strip(self, b: np.ndarray, dy: int, dx: int, halo: bool) -> np.ndarray:
        """View of the k-deep border strip facing (dy,dx) (halo=False) or the halo beyond it."""
        t, C, k = self.tile, self.grid.channels, self.k

        def span(d: int, n: int) -> slice:
            if d == 0:
                return slice(k, k + n)
            if halo:
                return slice(0, k) if d < 0 else slice(k + n, k + n + k)
            return slice(k, 2 * k) if d < 0 else slice(n, n + k)

        rs, cs = span(dy, t.rows), span(dx, t.cols)
        return b[rs, cs.start * C:cs.stop * C]

    def _post_exchange(self, b: np.ndarray):
        import torch
        dist = self.dist
        ops, recvs = [], []
        for (dy, dx), r in zip(NEIGHBOR_OFFSETS, self.nbrs):
            if r < 0 or (dy == 0 and dx == 0):
                continue
            if self.poison:
                self._strip(b, dy, dx, halo=True)[:] = 0xA5
            send = torch.from_numpy(np.ascontiguousarray(self._strip(b, dy, dx, halo=False)))
            view = self._strip(b, dy, dx, halo=True)
            recv = torch.empty(view.shape, dtype=torch.uint8)
            # tags keep the eight message streams apart: tag = direction as seen by the receiver
            d_send = (dy + 1) * 3 + (dx + 1)
            ops.append(dist.P2POp(dist.isend, send, r, tag=8 - d_send))
            ops.append(dist.P2POp(dist.irecv, recv, r, tag=d_send))
            recvs.append((view, recv))
        reqs = dist.batch_isend_irecv(ops) if ops else []
        return reqs, recvs

    @staticmethod
    def _finish_exchange(reqs, recvs) -> None:
        for q in reqs:
            q.wait()
        for view, recv in recvs:
            view[:] = recv.numpy()

    # -- iterations -----------------------------------------------------------------------------
    def _image_clip(self, r0: int, r1: int, c0: int, c1: int) -> tuple[int, int, int, int]:
        """Clip tile-relative coordinates to the image (outside stays zero = zero padding)."""
        t, g = self.tile, self.grid
        return (max(r0, -t.row0), min(r1, g.height - t.row0),
                max(c0, -t.col0), min(c1, g.width - t.col0))

    def _convolve(self, src: np.ndarray, dst: np.ndarray, r0: int, r1: int, c0: int, c1: int) -> None:
        r0, r1, c0, c1 = self._image_clip(r0, r1, c0, c1)
        if r0 < r1 and c0 < c1:
            H.cpu_convolve(src, dst, self.k, self.grid.channels, self.filt.spec, r0, r1, c0, c1,
                           self.threads)

    def iterate(self, loops: int) -> None:
        t, k = self.tile, self.k
        multi = self.grid.workers > 1 and self.dist is not None
        done = 0
        while done < loops:
            n = min(k, loops - done)
            src, dst = self.bufs[self.cur], self.bufs[self.cur ^ 1]
            if multi:
                reqs, recvs = self._post_exchange(src)
                self.exchanges += 1
                if n == 1:
                    # reference schedule: interior while the halos are in flight, then the ring
                    self._convolve(src, dst, 1, t.rows - 1, 1, t.cols - 1)
                    self._finish_exchange(reqs, recvs)
                    self._convolve(src, dst, 0, 1, 0, t.cols)
                    self._convolve(src, dst, t.rows - 1, t.rows, 0, t.cols)
                    self._convolve(src, dst, 1, t.rows - 1, 0, 1)
                    self._convolve(src, dst, 1, t.rows - 1, t.cols - 1, t.cols)
                    self.cur ^= 1
                    done += 1
                    continue
                self._finish_exchange(reqs, recvs)
            for j in range(1, n + 1):
                m = (n - j) if multi else 0   # recompute a shrinking ring of the halo
                self._convolve(src, dst, -m, t.rows + m, -m, t.cols + m)
                src, dst = dst, src
                self.cur ^= 1
            done += n


def run_single(image: np.ndarray, filt: Filter, loops: int, threads: int = 0) -> np.ndarray:
    """World-size-1 convenience path: whole image through the native CPU engine."""
    return H.cpu_run_image(image, filt.spec, loops, threads)
