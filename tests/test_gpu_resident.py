"""GPU numerics of the persistent resident kernel (in-kernel halo exchange between CTAs and between
virtual ranks) vs the oracle."""
import numpy as np
import pytest

from conftest import random_image
from parallel_convolution_b200.models.filters import Filter
from parallel_convolution_b200.ops import host

pytestmark = pytest.mark.gpu


def _run(img, filt, loops, workers=1, grid=None, halo_depth=None, band=None):
    from parallel_convolution_b200.parallel.decomposition import TileGrid
    from parallel_convolution_b200.parallel.fused import VirtualCluster
    h, w = img.shape[:2]
    c = 1 if img.ndim == 2 else img.shape[2]
    g = TileGrid.create(h, w, c, workers, grid)
    vc = VirtualCluster(g, filt, halo_depth=halo_depth, band=band, mode="resident", timeout_ms=4000.0)
    try:
        assert vc.rt.choose_mode(loops) == "resident"
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
def test_resident_single_rank(rng, name, c):
    img = random_image(rng, 300, 900, c)
    f = Filter.builtin(name)
    for loops, k in ((1, None), (7, 2), (12, 5), (9, 4)):
        ref = host.cpu_run_image(img, f.spec, loops)
        out = _run(img, f, loops, halo_depth=k)
        assert np.array_equal(out, ref), (name, c, loops, k, int((out != ref).sum()))


@pytest.mark.parametrize("band", [2, 4, 7, 14])
def test_resident_band_heights(rng, band):
    img = random_image(rng, 36 * band, 1200, 3)    # small enough to fit on chip with short bands
    f = Filter.builtin("gaussian")
    ref = host.cpu_run_image(img, f.spec, 10)
    assert np.array_equal(_run(img, f, 10, halo_depth=3, band=band), ref)


@pytest.mark.parametrize("workers,grid", [(2, None), (2, (1, 2)), (4, (2, 2)), (8, (4, 2)), (9, (3, 3))])
@pytest.mark.parametrize("c", [1, 3])
def test_resident_virtual_ranks(rng, workers, grid, c):
    img = random_image(rng, 360, 720, c)     # divisible by every grid above
    for name, loops, k in (("gaussian", 11, 3), ("sharpen", 6, 2)):
        f = Filter.builtin(name)
        ref = host.cpu_run_image(img, f.spec, loops)
        out = _run(img, f, loops, workers=workers, grid=grid, halo_depth=k)
        assert np.array_equal(out, ref), (workers, grid, c, name, int((out != ref).sum()))


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
    img = random_image(rng, 360, 720, 3)
    f = Filter.builtin("gaussian")
    g = TileGrid.create(360, 720, 3, 4, (2, 2))
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
