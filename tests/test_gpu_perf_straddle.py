"""Performance guard for the byte-wise slow path of the region I/O (region_io.cuh: load_chunk_bytes /
store_chunk_bytes): tiles whose width is not a whole number of 16-byte chunks make every chunk that straddles
a tile-ownership boundary go through one-byte loads. It must stay a boundary effect, not a cliff."""
import numpy as np
import pytest

from conftest import random_image
from parallel_convolution_b200.models.filters import Filter
from parallel_convolution_b200.ops import host

pytestmark = pytest.mark.gpu


def _time_ms(width, loops=20, reps=5):
    from parallel_convolution_b200.ops import cuda as K
    from parallel_convolution_b200.parallel.decomposition import TileGrid
    from parallel_convolution_b200.parallel.fused import VirtualCluster
    f = Filter.builtin("gaussian")
    g = TileGrid.create(600, width, 3, 2, (1, 2))
    vc = VirtualCluster(g, f, halo_depth=4, mode="stream", timeout_ms=4000.0)
    try:
        vc.synth(3)
        vc.run(loops)
        K.device_sync()
        best = None
        for _ in range(reps):
            e0, e1 = K.Event(), K.Event()
            e0.record(vc.rt.stream)
            vc.run(loops)
            e1.record(vc.rt.stream)
            e1.synchronize()
            ms = e0.elapsed_ms(e1)
            best = ms if best is None else min(best, ms)
        assert not vc.collect_status().watchdog
    finally:
        vc.close()
    return best


def test_odd_tile_widths_are_exact(rng):
    img = random_image(rng, 600, 1921, 3)
    from parallel_convolution_b200.parallel.decomposition import TileGrid
    from parallel_convolution_b200.parallel.fused import VirtualCluster
    f = Filter.builtin("gaussian")
    vc = VirtualCluster(TileGrid.create(600, 1921, 3, 2, (1, 2)), f, halo_depth=4, mode="stream")
    try:
        vc.upload(img)
        vc.run(9)
        out = vc.download().reshape(img.shape)
    finally:
        vc.close()
    assert np.array_equal(out, host.cpu_run_image(img, f.spec, 9))


def test_byte_wise_straddle_path_is_not_a_cliff():
    aligned = _time_ms(1920)      # 2 tiles of 960 px = 2880 bytes: every chunk has one owner
    odd = _time_ms(1921)          # tiles of 961 / 960 px: the boundary falls inside a chunk in every row
    print(f"straddle path: aligned {aligned:.4f} ms, odd {odd:.4f} ms, ratio {odd / aligned:.2f}")
    assert odd < 3.0 * aligned, (aligned, odd)
