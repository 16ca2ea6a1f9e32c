"""2-D block decomposition: tile grid, tile geometry, 8-neighbour table.

Reference behaviour reproduced (``mpi/mpi_conv.c``): ``divide_rows`` chooser (``:363-377``),
row-major rank grid (``:91-92``), N/S/W/E neighbours (``:150-157``). Extended by design:

* all **eight** neighbours — the reference never exchanges diagonal halo cells (defect D1,
  SURVEY.md §2.5), which makes its np>1 output differ from np=1;
* remainder tiles when the image is not divisible (the reference aborts, ``:57-61``);
* explicit ``grid=(rows, cols)`` override.

The arithmetic lives in C++ (``csrc/host/decomp.cpp``); this module is the typed Python face.
"""
from __future__ import annotations

from dataclasses import dataclass

from ..ops import host as _host

# (dy, dx) in the order used by every backend; index 4 is "self".
NEIGHBOR_OFFSETS = [(-1, -1), (-1, 0), (-1, 1), (0, -1), (0, 0), (0, 1), (1, -1), (1, 0), (1, 1)]
NEIGHBOR_NAMES = ["nw", "n", "ne", "w", "self", "e", "sw", "s", "se"]


@dataclass(frozen=True)
class Tile:
    rank: int
    ty: int
    tx: int
    row0: int
    rows: int
    col0: int   # in pixels
    cols: int   # in pixels


@dataclass(frozen=True)
class TileGrid:
    height: int
    width: int
    channels: int
    grid_rows: int
    grid_cols: int
    exact: bool

    @staticmethod
    def create(height: int, width: int, channels: int, workers: int,
               grid: tuple[int, int] | None = None) -> "TileGrid":
        if grid is not None:
            gr, gc = grid
            if gr * gc != workers:
                raise ValueError(f"grid {gr}x{gc} does not match {workers} workers")
            if gr > height or gc > width:
                raise ValueError("grid finer than the image")
            exact = height % gr == 0 and width % gc == 0
            return TileGrid(height, width, channels, gr, gc, exact)
        gr, gc, exact = _host.choose_grid(height, width, workers)
        return TileGrid(height, width, channels, gr, gc, exact)

    @property
    def workers(self) -> int:
        return self.grid_rows * self.grid_cols

    @property
    def row_bytes(self) -> int:
        return self.width * self.channels

    def coords(self, rank: int) -> tuple[int, int]:
        return rank // self.grid_cols, rank % self.grid_cols

    def rank_of(self, ty: int, tx: int) -> int:
        """Rank at grid position, or -1 outside the grid (the reference's ``-1`` neighbours)."""
        if 0 <= ty < self.grid_rows and 0 <= tx < self.grid_cols:
            return ty * self.grid_cols + tx
        return -1

    def tile(self, rank: int) -> Tile:
        ty, tx = self.coords(rank)
        r0, nr = _host.split(self.height, self.grid_rows, ty)
        c0, nc = _host.split(self.width, self.grid_cols, tx)
        return Tile(rank, ty, tx, r0, nr, c0, nc)

    def tiles(self) -> list[Tile]:
        return [self.tile(r) for r in range(self.workers)]

    def neighbors(self, rank: int) -> list[int]:
        """Nine entries in NEIGHBOR_OFFSETS order (index 4 = rank itself), -1 where absent."""
        ty, tx = self.coords(rank)
        return [self.rank_of(ty + dy, tx + dx) for dy, dx in NEIGHBOR_OFFSETS]

    def row_bounds(self) -> list[int]:
        """grid_rows + 1 global row boundaries."""
        b = [_host.split(self.height, self.grid_rows, i)[0] for i in range(self.grid_rows)]
        return b + [self.height]

    def col_bounds(self) -> list[int]:
        """grid_cols + 1 global pixel-column boundaries."""
        b = [_host.split(self.width, self.grid_cols, i)[0] for i in range(self.grid_cols)]
        return b + [self.width]

    def max_tile_rows(self) -> int:
        return -(-self.height // self.grid_rows)

    def max_tile_cols(self) -> int:
        return -(-self.width // self.grid_cols)

    def min_tile_rows(self) -> int:
        return self.height // self.grid_rows

    def min_tile_cols(self) -> int:
        return self.width // self.grid_cols

    def halo_bytes_per_iteration(self, rank: int, depth: int = 1) -> int:
        """Bytes a rank pulls from its neighbours per exchange with a ``depth``-pixel halo."""
        t = self.tile(rank)
        nb = self.neighbors(rank)
        C = self.channels
        total = 0
        for (dy, dx), r in zip(NEIGHBOR_OFFSETS, nb):
            if r < 0 or (dy == 0 and dx == 0):
                continue
            h = depth if dy != 0 else t.rows
            w = depth if dx != 0 else t.cols
            total += h * w * C
        return total

    def describe(self) -> dict:
        return {"grid": f"{self.grid_rows}x{self.grid_cols}", "exact": self.exact,
                "tile_rows": [self.min_tile_rows(), self.max_tile_rows()],
                "tile_cols": [self.min_tile_cols(), self.max_tile_cols()]}
