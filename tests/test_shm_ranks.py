"""CPU ranks over one shared image (csrc/host/shm_ranks.cpp, parallel/shm_backend.py).

The ranks of these tests are threads of one process attached to the same segment (the native call releases
the GIL), which exercises exactly the code separate processes run — the multi-process launch itself is
covered by tests/test_distributed_cpu.py. Everything is compared with the oracle byte for byte.
"""
import errno
import os
import threading
import time

import numpy as np
import pytest

from parallel_convolution_b200.models.filters import Filter
from parallel_convolution_b200.ops import host as H
from parallel_convolution_b200.parallel.decomposition import TileGrid
from parallel_convolution_b200.parallel.shm_backend import ShmTile


def _segment(h, w, c, workers):
    path = f"/dev/shm/pconv_test_{os.getpid()}_{threading.get_ident()}_{time.monotonic_ns()}"
    seg = H.ShmSegment(path, h, w, c, workers, create=True)
    os.unlink(path)                      # mapped: the name is not needed any more
    return seg


def _ranks(h, w, c, workers, grid, filt, k, threads=1, timeout_ms=20_000.0):
    g = TileGrid.create(h, w, c, workers, grid)
    f = Filter.builtin(filt) if isinstance(filt, str) else filt
    seg = _segment(h, w, c, workers)
    return g, f, seg, [ShmTile(g, r, f, k, threads, None, timeout_ms, segment=seg) for r in range(workers)]


def _load(tiles, img2d, c):
    for t in tiles:
        tt = t.tile
        t.seg.reset(t.rank)
        t._view(0)[:] = img2d[tt.row0:tt.row0 + tt.rows, tt.col0 * c:(tt.col0 + tt.cols) * c]


def _run_all(tiles, loops, calls=1):
    errs = []

    def work(t):
        try:
            for _ in range(calls):
                t.iterate(loops)
        except Exception as e:  # noqa: BLE001
            errs.append((t.rank, e))

    th = [threading.Thread(target=work, args=(t,)) for t in tiles]
    [x.start() for x in th]
    [x.join() for x in th]
    return errs


def _gather(tiles, h, rb, c):
    out = np.empty((h, rb), dtype=np.uint8)
    for t in tiles:
        tt = t.tile
        out[tt.row0:tt.row0 + tt.rows, tt.col0 * c:(tt.col0 + tt.cols) * c] = t.result()
    return out


@pytest.mark.parametrize("h,w,c,workers,grid,filt,loops,k,threads,calls", [
    (64, 96, 1, 4, None, "gaussian", 7, 3, 1, 1),        # 2x2: diagonal neighbours
    (64, 96, 3, 4, None, "gaussian", 10, 4, 1, 2),       # epochs continue across calls; last epoch shorter
    (61, 97, 3, 6, None, "sharpen", 9, 2, 1, 1),         # remainder tiles, negative weights
    (50, 40, 4, 8, None, "box", 5, 5, 1, 1),             # exact-fp32 path, k == loops
    (120, 64, 1, 3, (3, 1), "gaussian", 13, 6, 2, 1),    # rows only: interior bands wait for nobody; 2 threads
    (64, 120, 3, 3, (1, 3), "edge", 11, 6, 1, 1),        # columns only
    (200, 300, 3, 9, (3, 3), "gaussian", 12, 1, 1, 1),   # k = 1: straight from image to image
    (37, 29, 1, 1, None, "gaussian", 9, 4, 3, 1),        # one rank, more threads than bands
])
def test_thread_ranks_match_the_oracle(h, w, c, workers, grid, filt, loops, k, threads, calls):
    if os.environ.get("PCONV_TEST_SINGLE_THREAD_RANKS"):
        threads = 1      # ThreadSanitizer run (scripts/host_sanitize.sh): libgomp's barriers are not instrumented
    g, f, seg, tiles = _ranks(h, w, c, workers, grid, filt, k, threads)
    shape = (h, w) if c == 1 else (h, w, c)
    img = np.random.default_rng(h * w + workers).integers(0, 256, shape, dtype=np.uint8)
    _load(tiles, img.reshape(h, w * c), c)
    assert not _run_all(tiles, loops, calls)
    want = H.oracle(img, f.spec, loops * calls).reshape(h, w * c)
    assert np.array_equal(_gather(tiles, h, w * c, c), want)
    epochs = -(-loops // k) * calls
    assert all(t.exchanges == epochs and seg.epoch(t.rank) == epochs for t in tiles)
    seg.close()


def test_reload_starts_a_new_run():
    g, f, seg, tiles = _ranks(48, 40, 1, 4, None, "gaussian", 3)
    rng = np.random.default_rng(9)
    for _ in range(2):
        img = rng.integers(0, 256, (48, 40), dtype=np.uint8)
        _load(tiles, img, 1)
        assert not _run_all(tiles, 5)                    # 2 epochs: the result sits in image 0 again
        assert np.array_equal(_gather(tiles, 48, 40, 1), H.oracle(img, f.spec, 5).reshape(48, 40))
    seg.close()


def test_a_stalled_rank_is_detected_not_waited_for_forever():
    """Rank 3 never runs: its neighbours time out (ETIMEDOUT) and raise their abort flags, so the ranks
    that wait on *them* stop with ECANCELED instead of sitting out their own timeout."""
    g, f, seg, tiles = _ranks(64, 64, 1, 4, (4, 1), "gaussian", 2, timeout_ms=300.0)
    img = np.random.default_rng(1).integers(0, 256, (64, 64), dtype=np.uint8)
    _load(tiles, img, 1)
    tiles[3].inject_stall = True
    t0 = time.perf_counter()
    errs = dict(_run_all(tiles, 12))
    assert time.perf_counter() - t0 < 10.0
    assert 3 not in errs and set(errs) == {0, 1, 2}
    codes = {r: e.errno for r, e in errs.items()}
    assert codes[2] == errno.ETIMEDOUT                   # the direct neighbour of the stalled rank
    assert all(c in (errno.ETIMEDOUT, errno.ECANCELED) for c in codes.values())
    seg.close()


def test_rejects_bad_geometry():
    seg = _segment(32, 32, 1, 2)
    f = Filter.builtin("gaussian")
    with pytest.raises(OSError):
        seg.run(0, (0, 40, 0, 32), [-1] * 9, f.spec, 1, 1)          # tile taller than the image
    with pytest.raises(OSError):
        seg.run(5, (0, 16, 0, 32), [-1] * 9, f.spec, 1, 1)          # rank outside the segment
    with pytest.raises(OSError):
        H.ShmSegment("/dev/shm/pconv_definitely_missing_segment")    # nothing to attach to
    seg.close()
    g = TileGrid.create(32, 32, 1, 4, (4, 1))
    with pytest.raises(ValueError):
        ShmTile(g, 0, f, 9, 1, None, segment=_segment(32, 32, 1, 4))  # halo deeper than a tile (8 rows)


def test_band_helper_equals_full_image_passes():
    """pconv_cpu_run_image_blocked (bands + k-deep halos in private buffers) == plain passes == oracle."""
    img = np.random.default_rng(2).integers(0, 256, (300, 1500, 3), dtype=np.uint8)   # > 1 band per thread
    f = Filter.builtin("gaussian")
    want = H.oracle(img, f.spec, 7)
    for k, threads in ((1, 1), (3, 2), (7, 4), (5, 1)):
        assert np.array_equal(H.cpu_run_image_blocked(img, f.spec, 7, k, threads), want)
