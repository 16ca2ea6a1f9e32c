"""GPU numerics of the persistent kernel (persist_kernel.cuh: pushed in-kernel halo exchange between CTAs
and between virtual ranks, bulk-copy inbox on an mbarrier) vs the oracle."""
import numpy as np
import pytest

from conftest import random_image
from parallel_convolution_b200.models.filters import Filter
from parallel_convolution_b200.ops import host

pytestmark = pytest.mark.gpu


def _run(img, filt, loops, workers=1, grid=None, halo_depth=None, band=None, warps=None):
    from parallel_convolution_b200.parallel.decomposition import TileGrid
    from parallel_convolution_b200.parallel.fused import VirtualCluster
    h, w = img.shape[:2]
    c = 1 if img.ndim == 2 else img.shape[2]
    g = TileGrid.create(h, w, c, workers, grid)
    vc = VirtualCluster(g, filt, halo_depth=halo_depth, band=band, mode="resident", timeout_ms=4000.0,
                        warps=warps)
    try:
        assert vc.rt.choose_mode(loops) == "resident"
        if warps:
            assert vc.rt.resident_plan(loops).W == warps
        vc.upload(img)
        vc.run(loops)
        out = vc.download().reshape(img.shape)
        st = vc.collect_status()
        assert not st.watchdog, "halo watchdog fired (CTAs / virtual ranks not co-resident?)"
    finally:
        vc.close()
    return out


@pytest.mark.parametrize("name", ["gaussian", "box", "sharpen", "edge"])
@pytest.mark.parametrize("c", [1, 3])
@pytest.mark.parametrize("warps", [8, 16])
def test_resident_single_rank(rng, name, c, warps):
    img = random_image(rng, 300, 900, c)
    f = Filter.builtin(name)
    for loops, k in ((1, None), (7, 2), (12, 5), (9, 4)):
        ref = host.cpu_run_image(img, f.spec, loops)
        out = _run(img, f, loops, halo_depth=k, warps=warps)
        assert np.array_equal(out, ref), (name, c, loops, k, warps, int((out != ref).sum()))


@pytest.mark.parametrize("band", [2, 3, 5, 7, 12, 14])
@pytest.mark.parametrize("warps", [8, 16])
def test_resident_band_heights(rng, band, warps):
    img = random_image(rng, 36 * band, 1200, 3)    # small enough to fit on chip with short bands
    f = Filter.builtin("gaussian")
    ref = host.cpu_run_image(img, f.spec, 10)
    assert np.array_equal(_run(img, f, 10, halo_depth=3, band=band, warps=warps), ref)


@pytest.mark.parametrize("h,w,c", [(211, 997, 1), (130, 1003, 3), (97, 333, 4), (40, 64, 1), (500, 171, 3)])
def test_resident_odd_shapes(rng, h, w, c):
    """Ragged last CTA row and column (partial chunks at the image edge), one rank."""
    img = random_image(rng, h, w, c)
    for name, loops, k in (("gaussian", 9, 2), ("emboss", 5, 3)):
        f = Filter.builtin(name)
        ref = host.cpu_run_image(img, f.spec, loops)
        for warps in (8, 16):
            out = _run(img, f, loops, halo_depth=k, warps=warps)
            assert np.array_equal(out, ref), (h, w, c, name, warps, int((out != ref).sum()))


@pytest.mark.parametrize("workers,grid", [(2, None), (2, (1, 2)), (4, (2, 2)), (8, (4, 2)), (9, (3, 3))])
@pytest.mark.parametrize("c", [1, 3])
def test_resident_virtual_ranks(rng, workers, grid, c):
    img = random_image(rng, 360, 768, c)     # divisible by every grid above into whole 16-byte chunks
    for name, loops, k, warps in (("gaussian", 11, 3, 8), ("sharpen", 6, 2, 16)):
        f = Filter.builtin(name)
        ref = host.cpu_run_image(img, f.spec, loops)
        out = _run(img, f, loops, workers=workers, grid=grid, halo_depth=k, warps=warps)
        assert np.array_equal(out, ref), (workers, grid, c, name, int((out != ref).sum()))


def test_unsupported_geometry_falls_back_to_stream(rng):
    """Tiles that are not whole 16-byte chunks wide cannot use the pushed exchange across ranks: `auto`
    streams, an explicit `resident` request says why."""
    from parallel_convolution_b200.parallel.decomposition import TileGrid
    from parallel_convolution_b200.parallel.fused import VirtualCluster
    img = random_image(rng, 120, 362, 3)                 # 2 tile columns of 181 px = 543 bytes
    f = Filter.builtin("gaussian")
    g = TileGrid.create(120, 362, 3, 2, (1, 2))
    vc = VirtualCluster(g, f, mode="auto", timeout_ms=4000.0)
    try:
        assert vc.rt.choose_mode(6) == "stream"
        vc.upload(img)
        vc.run(6)
        assert np.array_equal(vc.download().reshape(img.shape), host.cpu_run_image(img, f.spec, 6))
    finally:
        vc.close()
    vc = VirtualCluster(g, f, mode="resident", timeout_ms=4000.0)
    try:
        with pytest.raises(ValueError):
            vc.rt.choose_mode(6)
    finally:
        vc.close()


def test_resident_reference_config(rng):
    img = random_image(rng, 2520, 1920, 3)
    f = Filter.builtin("gaussian")
    ref = host.cpu_run_image(img, f.spec, 23)
    assert np.array_equal(_run(img, f, 23), ref)
    assert np.array_equal(_run(img, f, 23, workers=2), ref)


def test_resident_back_to_back_runs(rng):
    """Epoch continuity: two runs on the same runtime without re-creating anything."""
    from parallel_convolution_b200.parallel.decomposition import TileGrid
    from parallel_convolution_b200.parallel.fused import VirtualCluster
    img = random_image(rng, 360, 768, 3)
    f = Filter.builtin("gaussian")
    g = TileGrid.create(360, 768, 3, 4, (2, 2))
    vc = VirtualCluster(g, f, halo_depth=3, mode="resident", timeout_ms=4000.0)
    try:
        vc.upload(img)
        vc.run(7)
        vc.run(5)   # continues from the previous result
        out = vc.download().reshape(img.shape)
        assert not vc.collect_status().watchdog
    finally:
        vc.close()
    assert np.array_equal(out, host.cpu_run_image(img, f.spec, 12))
