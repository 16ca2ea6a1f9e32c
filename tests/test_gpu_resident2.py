"""GPU numerics of the split-phase resident kernel (two virtual tiles per CTA, halo refresh of the
idle tile overlapped with the arithmetic of the other — csrc/cuda/resident2_kernel.cuh) vs the oracle."""
import numpy as np
import pytest

from conftest import random_image
from parallel_convolution_b200.models.filters import Filter
from parallel_convolution_b200.ops import host

pytestmark = pytest.mark.gpu


def _run(img, filt, loops, workers=1, grid=None, halo_depth=None, band=None, runs=None):
    from parallel_convolution_b200.parallel.decomposition import TileGrid
    from parallel_convolution_b200.parallel.fused import VirtualCluster
    h, w = img.shape[:2]
    c = 1 if img.ndim == 2 else img.shape[2]
    g = TileGrid.create(h, w, c, workers, grid)
    vc = VirtualCluster(g, filt, halo_depth=halo_depth, band=band, mode="resident2", timeout_ms=4000.0)
    try:
        assert vc.rt.choose_mode(loops) == "resident2"
        plan = vc.rt.plan_for("resident2", loops)
        assert plan.dual and plan.nty % 2 == 0 and plan.nty // 2 * plan.ntx <= vc.rt.sms
        vc.upload(img)
        for n in (runs or [loops]):
            vc.run(n)
        out = vc.download().reshape(img.shape)
        st = vc.collect_status()
        assert not st.watchdog, "halo watchdog fired"
    finally:
        vc.close()
    return out


@pytest.mark.parametrize("name", ["gaussian", "box", "sharpen", "edge"])
@pytest.mark.parametrize("c", [1, 3])
def test_resident2_single_rank(rng, name, c):
    img = random_image(rng, 300, 900, c)
    f = Filter.builtin(name)
    for loops, k in ((1, None), (7, 2), (12, 5), (9, 4), (10, 1)):
        ref = host.cpu_run_image(img, f.spec, loops)
        out = _run(img, f, loops, halo_depth=k)
        assert np.array_equal(out, ref), (name, c, loops, k, int((out != ref).sum()))


@pytest.mark.parametrize("band", [2, 4, 5, 7])
def test_resident2_band_heights(rng, band):
    img = random_image(rng, 60 * band, 1200, 3)
    f = Filter.builtin("gaussian")
    ref = host.cpu_run_image(img, f.spec, 10)
    assert np.array_equal(_run(img, f, 10, halo_depth=3, band=band), ref)


def test_resident2_odd_shapes(rng):
    """Tile heights/widths that are no multiple of anything: partial and empty virtual tiles, chunks
    straddling the tile edge, RGBA."""
    for h, w, c, k in ((301, 333, 3, 3), (97, 1001, 1, 2), (500, 257, 4, 4)):
        img = random_image(rng, h, w, c)
        f = Filter.builtin("gaussian" if c != 1 else "laplacian")
        ref = host.cpu_run_image(img, f.spec, 9)
        out = _run(img, f, 9, halo_depth=k)
        assert np.array_equal(out, ref), (h, w, c, int((out != ref).sum()))


@pytest.mark.parametrize("workers,grid", [(2, None), (2, (1, 2)), (4, (2, 2)), (8, (4, 2)), (9, (3, 3))])
@pytest.mark.parametrize("c", [1, 3])
def test_resident2_virtual_ranks(rng, workers, grid, c):
    img = random_image(rng, 360, 720, c)     # divisible by every grid above
    for name, loops, k in (("gaussian", 11, 3), ("sharpen", 6, 2)):
        f = Filter.builtin(name)
        ref = host.cpu_run_image(img, f.spec, loops)
        out = _run(img, f, loops, workers=workers, grid=grid, halo_depth=k)
        assert np.array_equal(out, ref), (workers, grid, c, name, int((out != ref).sum()))


def test_resident2_reference_config(rng):
    img = random_image(rng, 2520, 1920, 3)
    f = Filter.builtin("gaussian")
    ref = host.cpu_run_image(img, f.spec, 23)
    assert np.array_equal(_run(img, f, 23), ref)
    assert np.array_equal(_run(img, f, 23, workers=2), ref)


def test_resident2_back_to_back_runs(rng):
    """Epoch continuity: two runs on the same runtime without re-creating anything."""
    img = random_image(rng, 360, 720, 3)
    f = Filter.builtin("gaussian")
    out = _run(img, f, 7, workers=4, grid=(2, 2), halo_depth=3, runs=[7, 5])
    assert np.array_equal(out, host.cpu_run_image(img, f.spec, 12))
