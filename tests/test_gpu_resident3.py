"""Opt-in GPU tests of the split-phase resident kernel with a protocol warp (resident3_kernel.cuh).

The kernel synchronises its compute warps and its protocol warp with NAMED barriers; a mistake there
hangs the kernel where no watchdog can reach it. These tests passed on a B200 (scripts/r3_bringup.sh,
profiles/final_r1/resident3_pytest_tail.txt) but stay opt-in while the kernel is experimental: ``PCONV_TEST_RESIDENT3=1 timeout 120 python -m pytest tests/test_gpu_resident3.py``.
"""
import os

import numpy as np
import pytest

from conftest import random_image
from parallel_convolution_b200.models.filters import Filter
from parallel_convolution_b200.ops import host

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(os.environ.get("PCONV_TEST_RESIDENT3") != "1",
                                 reason="experimental kernel: set PCONV_TEST_RESIDENT3=1 (run under `timeout`)")]


def _run(img, filt, loops, workers=1, grid=None, halo_depth=None, band=None):
    from parallel_convolution_b200.parallel.decomposition import TileGrid
    from parallel_convolution_b200.parallel.fused import VirtualCluster
    h, w = img.shape[:2]
    c = 1 if img.ndim == 2 else img.shape[2]
    g = TileGrid.create(h, w, c, workers, grid)
    vc = VirtualCluster(g, filt, halo_depth=halo_depth, band=band, mode="resident3", timeout_ms=4000.0)
    try:
        assert vc.rt.choose_mode(loops) == "resident3"
        vc.upload(img)
        vc.run(loops)
        out = vc.download().reshape(img.shape)
        assert not vc.collect_status().watchdog, "halo watchdog fired"
    finally:
        vc.close()
    return out


@pytest.mark.parametrize("name,c", [("gaussian", 3), ("sharpen", 1), ("box", 3)])
def test_resident3_single_rank(rng, name, c):
    img = random_image(rng, 300, 900, c)
    f = Filter.builtin(name)
    for loops, k in ((1, None), (7, 2), (12, 5), (10, 1)):
        ref = host.cpu_run_image(img, f.spec, loops)
        assert np.array_equal(_run(img, f, loops, halo_depth=k), ref), (name, c, loops, k)


@pytest.mark.parametrize("workers,grid", [(2, None), (4, (2, 2)), (9, (3, 3))])
def test_resident3_virtual_ranks(rng, workers, grid):
    img = random_image(rng, 360, 720, 3)
    f = Filter.builtin("gaussian")
    ref = host.cpu_run_image(img, f.spec, 11)
    assert np.array_equal(_run(img, f, 11, workers=workers, grid=grid, halo_depth=3), ref)
