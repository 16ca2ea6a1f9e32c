"""GPU numerics: the streaming sm_100a kernel vs the fp32 oracle, byte for byte.

Covers every arithmetic path (sep121 / int / f32) x grey / RGB, temporal-blocking depths,
iteration parity (reference defect D4), image sizes that are not multiples of anything, and
decomposition invariance over virtual ranks on one device (the test the reference's np>1 runs
would fail: defects D1/D2).
"""
import numpy as np
import pytest

from conftest import random_image
from parallel_convolution_b200.models.filters import Filter
from parallel_convolution_b200.ops import host

pytestmark = pytest.mark.gpu


def _cluster(h, w, c, workers, filt, grid=None, halo_depth=None):
    from parallel_convolution_b200.parallel.decomposition import TileGrid
    from parallel_convolution_b200.parallel.fused import VirtualCluster
    g = TileGrid.create(h, w, c, workers, grid)
    return VirtualCluster(g, filt, halo_depth=halo_depth)


def _run(img, filt, loops, workers=1, grid=None, halo_depth=None):
    h, w = img.shape[:2]
    c = 1 if img.ndim == 2 else img.shape[2]
    vc = _cluster(h, w, c, workers, filt, grid, halo_depth)
    try:
        vc.upload(img)
        vc.run(loops)
        out = vc.download().reshape(img.shape)
        st = vc.collect_status()
        assert not st.watchdog
    finally:
        vc.close()
    return out


@pytest.mark.parametrize("name", ["gaussian", "box", "edge", "sharpen", "laplacian", "emboss",
                                  "sobel_x", "identity", "gaussian_soft"])
@pytest.mark.parametrize("c", [1, 3])
def test_filters_match_oracle(rng, name, c):
    img = random_image(rng, 97, 131, c)
    f = Filter.builtin(name)
    for loops in (1, 2, 5):
        ref = host.oracle(img, f.spec, loops)
        out = _run(img, f, loops)
        assert np.array_equal(out, ref), f"{name} c={c} loops={loops}: {(out != ref).sum()} bytes differ"


@pytest.mark.parametrize("c", [1, 3])
@pytest.mark.parametrize("k", [1, 2, 3, 7])
def test_temporal_blocking_depths(rng, c, k):
    img = random_image(rng, 300, 700, c)
    f = Filter.builtin("gaussian")
    ref = host.oracle(img, f.spec, 9)
    out = _run(img, f, 9, halo_depth=k)
    assert np.array_equal(out, ref)


@pytest.mark.parametrize("loops", [0, 1, 2, 3, 10, 11])
def test_iteration_parity(rng, loops):
    img = random_image(rng, 64, 80, 3)
    f = Filter.builtin("gaussian")
    assert np.array_equal(_run(img, f, loops), host.oracle(img, f.spec, loops))


@pytest.mark.parametrize("shape", [(1, 1), (1, 40), (40, 1), (2, 3), (17, 511), (513, 19), (250, 1000)])
@pytest.mark.parametrize("c", [1, 3])
def test_odd_shapes(rng, shape, c):
    img = random_image(rng, shape[0], shape[1], c)
    for name in ("gaussian", "sharpen", "box"):
        f = Filter.builtin(name)
        assert np.array_equal(_run(img, f, 3), host.oracle(img, f.spec, 3)), (shape, c, name)


@pytest.mark.parametrize("workers,grid", [(2, None), (2, (1, 2)), (4, None), (8, None), (6, (3, 2)),
                                          (9, (3, 3)), (8, (2, 4))])
@pytest.mark.parametrize("c", [1, 3])
def test_decomposition_invariance(rng, workers, grid, c):
    img = random_image(rng, 240, 336, c)
    for name, loops in (("gaussian", 7), ("edge", 3), ("sharpen", 4)):
        f = Filter.builtin(name)
        ref = host.oracle(img, f.spec, loops)
        out = _run(img, f, loops, workers=workers, grid=grid)
        assert np.array_equal(out, ref), (workers, grid, c, name)


@pytest.mark.parametrize("h,w,c,grid,k", [(40, 64, 1, (4, 1), 4), (48, 2048, 3, (8, 1), 4), (64, 300, 1, (1, 4), 3),
                                          (36, 96, 3, (3, 3), 2)])
def test_small_tiles_in_deep_grids(rng, h, w, c, grid, k):
    """Tiles much smaller than a CTA region in grids three or more tiles deep: the region overhangs the
    ADJACENT tile, and what lies beyond it belongs to a tile two steps away — it must be treated as zero,
    never read (it would fall outside that neighbour's buffer). Run under compute-sanitizer memcheck by
    scripts/sanitize.sh."""
    img = random_image(rng, h, w, c)
    f = Filter.builtin("gaussian")
    out = _run(img, f, 2 * k + 1, workers=grid[0] * grid[1], grid=grid, halo_depth=k)
    assert np.array_equal(out, host.cpu_run_image(img, f.spec, 2 * k + 1))


def test_remainder_tiles(rng):
    img = random_image(rng, 101, 203, 3)   # not divisible by anything useful
    f = Filter.builtin("gaussian")
    ref = host.oracle(img, f.spec, 6)
    for workers, grid in ((4, (2, 2)), (6, (2, 3)), (8, (4, 2))):
        assert np.array_equal(_run(img, f, 6, workers=workers, grid=grid), ref), (workers, grid)


def test_custom_float_filter(rng):
    img = random_image(rng, 90, 120, 3)
    f = Filter.custom([0.05, 0.1, 0.05, 0.1, 0.4, 0.1, 0.05, 0.1, 0.05])
    assert f.kind_name == "f32-exact"
    assert np.array_equal(_run(img, f, 4), host.oracle(img, f.spec, 4))
    g = Filter.custom([-0.3, 0.7, -0.3, 0.9, 1.3, 0.9, -0.3, 0.7, -0.3])  # overshoots both ways
    assert np.array_equal(_run(img, g, 3), host.oracle(img, g.spec, 3))


def test_reference_image_config(rng):
    """BASELINE config shape (1920x2520 RGB), fewer iterations to keep the oracle affordable."""
    img = random_image(rng, 2520, 1920, 3)
    f = Filter.builtin("gaussian")
    ref = host.cpu_run_image(img, f.spec, 12)   # CPU engine == oracle (tests/test_cpu_engine.py)
    assert np.array_equal(_run(img, f, 12), ref)
    assert np.array_equal(_run(img, f, 12, workers=8), ref)


@pytest.mark.parametrize("name", ["gaussian", "sharpen", "box"])
def test_rgba_four_channel_images(rng, name):
    """Channel stride 4 (RGBA): an extension over the reference's grey / rgb."""
    img = random_image(rng, 120, 200, 4)
    f = Filter.builtin(name)
    ref = host.oracle(img, f.spec, 5)
    assert np.array_equal(_run(img, f, 5), ref)
    assert np.array_equal(_run(img, f, 5, workers=4), ref)


def test_device_synth_matches_host():
    from parallel_convolution_b200.parallel.decomposition import TileGrid
    from parallel_convolution_b200.parallel.fused import VirtualCluster
    g = TileGrid.create(50, 70, 3, 4)
    vc = VirtualCluster(g, Filter.builtin("gaussian"))
    try:
        vc.synth(99)
        dev = vc.download()
    finally:
        vc.close()
    assert np.array_equal(dev, host.synth_rect(99, 210, 0, 50, 0, 210))
