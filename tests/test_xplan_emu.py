"""The persistent kernel's pushed halo exchange, replayed on the CPU with the device code's own index
functions (csrc/common/xplan.h, csrc/host/xplan_emu.cpp) and compared byte for byte with the oracle.

Covers what a 1-GPU box cannot: ragged last CTA rows / columns, all eight neighbour directions, the
wrap of CTA coordinates into the neighbouring rank, both warp counts and odd epochs (inbox parity)."""
import numpy as np
import pytest

from parallel_convolution_b200.models.filters import Filter
from parallel_convolution_b200.ops import host as H

GAUSS = Filter.parse("gaussian")
BOX = Filter.parse("box")


def _img(h, w, c, seed=0):
    rng = np.random.default_rng(seed)
    shape = (h, w) if c == 1 else (h, w, c)
    return rng.integers(0, 256, shape, dtype=np.uint8)


CASES = [
    # H,   W,   C, grid,   k, R, Wp, TH, loops
    (96,  700, 1, (1, 1), 3, 4, 8, 0, 7),      # one rank, several CTA columns (ragged last column)
    (150, 400, 3, (1, 1), 2, 2, 16, 0, 5),     # RGB, W = 16
    (64,  320, 1, (2, 2), 2, 4, 8, 0, 6),      # 2x2 ranks, all eight directions cross ranks
    (90,  256, 3, (3, 1), 3, 2, 8, 0, 7),      # three ranks stacked (N and S neighbours both remote)
    (60,  512, 1, (1, 4), 2, 4, 8, 0, 5),      # four ranks side by side
    (120, 640, 1, (2, 2), 5, 3, 8, 9, 11),     # explicit TH -> ragged last CTA row, odd loops
    (40,  1600, 1, (1, 1), 16, 7, 8, 0, 33),   # HB = 16 exactly at k = 16 (grey)
    (48,  200, 3, (1, 1), 6, 4, 8, 0, 12),     # RGB k = 6 -> HB = 32 (two halo chunks per side)
    (48,  400, 4, (2, 1), 3, 4, 16, 0, 6),     # RGBA
]


@pytest.mark.parametrize("h,w,c,grid,k,R,Wp,TH,loops", CASES)
def test_emulated_exchange_matches_oracle(h, w, c, grid, k, R, Wp, TH, loops):
    img = _img(h, w, c, seed=h + w)
    for epoch_base in (0, 7):
        got = H.xplan_emulate(img, GAUSS.spec, loops, grid, k, R, Wp, TH, epoch_base)
        assert not isinstance(got, int), f"geometry rejected: {got}"
        want = H.oracle(img, GAUSS.spec, loops)
        assert np.array_equal(got, want)


def test_emulated_exchange_generic_filter():
    img = _img(70, 300, 3, seed=5)
    got = H.xplan_emulate(img, BOX.spec, 5, (2, 1), 2, 4, 8)
    assert np.array_equal(got, H.oracle(img, BOX.spec, 5))


def test_rejected_geometries():
    img = _img(64, 100, 3)
    # 100 px RGB over 2 ranks: 150-byte tiles are not whole chunks
    assert H.xplan_emulate(img, GAUSS.spec, 4, (1, 2), 2, 4, 8) == -5
    # halo deeper than the band allows
    assert H.xplan_emulate(img, GAUSS.spec, 4, (1, 1), 9, 2, 8) == -2
