"""CPU tests: oracle vs an independent NumPy model, CPU engine vs oracle, decomposition, I/O."""
import os

import numpy as np
import pytest

from conftest import random_image
from parallel_convolution_b200.models.filters import BUILTIN, Filter
from parallel_convolution_b200.ops import host
from parallel_convolution_b200.parallel.decomposition import TileGrid


def numpy_oracle(img, taps, loops):
    """Independent model of the numerics contract (fp32, tap order, separate mul/add roundings)."""
    a = img.astype(np.uint8)
    if a.ndim == 2:
        a = a[:, :, None]
    H, W, C = a.shape
    for _ in range(loops):
        p = np.zeros((H + 2, W + 2, C), dtype=np.float32)
        p[1:-1, 1:-1] = a
        val = np.zeros((H, W, C), dtype=np.float32)
        for k in range(3):
            for l in range(3):
                prod = (p[k:k + H, l:l + W] * np.float32(taps[k, l])).astype(np.float32)
                val = (val + prod).astype(np.float32)
        a = np.clip(np.trunc(val), 0, 255).astype(np.uint8)
    return a.reshape(img.shape)


@pytest.mark.parametrize("name", sorted(BUILTIN))
@pytest.mark.parametrize("c", [1, 3])
def test_oracle_matches_numpy_model(rng, name, c):
    img = random_image(rng, 23, 31, c)
    f = Filter.builtin(name)
    assert np.array_equal(host.oracle(img, f.spec, 3), numpy_oracle(img, f.taps, 3))


@pytest.mark.parametrize("name", sorted(BUILTIN))
@pytest.mark.parametrize("c", [1, 3, 4])
def test_cpu_engine_matches_oracle(rng, name, c):
    img = random_image(rng, 41, 67, c)
    f = Filter.builtin(name)
    for loops in (0, 1, 4):
        assert np.array_equal(host.cpu_run_image(img, f.spec, loops, threads=3),
                              host.oracle(img, f.spec, loops))


def test_gaussian_is_integer_shift(rng):
    """Gaussian/16 in fp32 == (sum w*p) >> 4 for every input."""
    img = random_image(rng, 64, 64, 1)
    f = Filter.builtin("gaussian")
    p = np.zeros((66, 66), dtype=np.int64)
    p[1:-1, 1:-1] = img
    w = f.spec.weights()
    s = sum(w[k, l] * p[k:k + 64, l:l + 64] for k in range(3) for l in range(3))
    assert np.array_equal(host.oracle(img, f.spec, 1), (s >> 4).astype(np.uint8))


def test_filter_classification():
    kinds = {n: Filter.builtin(n).kind_name for n in BUILTIN}
    assert kinds["gaussian"] == "sep121-swar"
    assert kinds["box"] == "f32-exact" and kinds["edge"] == "f32-exact"
    assert kinds["sharpen"] == "int-swar" and kinds["laplacian"] == "int-swar"
    assert Filter.parse("1,2,1,2,4,2,1,2,1/16").kind_name == "sep121-swar"
    assert Filter.parse("blur").name == "gaussian"
    assert Filter.custom([0.25, 0, 0, 0, 0.5, 0, 0, 0, 0.25]).kind_name == "int-swar"
    assert Filter.from_int([100] * 9, 1).kind_name == "f32-exact"   # sum|w| too large for SWAR


def test_divide_rows_matches_reference_decisions():
    # SURVEY.md §3.5, verified against mpi/mpi_conv.c:363-377
    assert [host.divide_rows(2520, 1920, n) for n in (1, 2, 4, 8, 9, 16, 25)] == [1, 2, 2, 4, 3, 4, 5]
    assert host.divide_rows(16384, 16384, 8) == 2       # tie -> fewer tile rows: 2x4
    assert host.divide_rows(7, 7, 4) == 0               # the reference aborts here
    g = TileGrid.create(7, 7, 1, 4)                     # we fall back to remainder tiles
    assert (g.grid_rows, g.grid_cols, g.exact) == (2, 2, False)
    assert sum(t.rows * t.cols for t in g.tiles()) == 49


def test_tile_grid_covers_image_and_neighbors():
    g = TileGrid.create(101, 203, 3, 6, (2, 3))
    cover = np.zeros((101, 203), dtype=int)
    for t in g.tiles():
        cover[t.row0:t.row0 + t.rows, t.col0:t.col0 + t.cols] += 1
    assert (cover == 1).all()
    nb = g.neighbors(4)   # centre-right of a 2x3 grid: ty=1, tx=1
    assert nb == [0, 1, 2, 3, 4, 5, -1, -1, -1]
    assert g.neighbors(0) == [-1, -1, -1, -1, 0, 1, -1, 3, 4]


def test_raw_io_roundtrip(tmp_path, rng):
    img = random_image(rng, 33, 45, 3)
    path = os.fspath(tmp_path / "img.raw")
    img.tofile(path)
    assert host.file_size(path) == img.size
    rb = 45 * 3
    rect = host.read_rect(path, rb, 5, 10, 6, 30)
    assert np.array_equal(rect, img.reshape(33, rb)[5:15, 6:36])
    out = os.fspath(tmp_path / "out.raw")
    host.create_file(out, img.size)
    g = TileGrid.create(33, 45, 3, 4, (2, 2))
    for t in g.tiles():   # tiles land in any order, like the rank-parallel write
        host.write_rect(out, rb, t.row0, t.col0 * 3,
                        np.ascontiguousarray(img.reshape(33, rb)[t.row0:t.row0 + t.rows,
                                                                 t.col0 * 3:(t.col0 + t.cols) * 3]))
    assert np.array_equal(np.fromfile(out, dtype=np.uint8), img.reshape(-1))
    with pytest.raises(OSError):
        host.read_rect(path, rb, 30, 10, 0, rb)   # short file must fail, not spin (cuda/funcs.c:39)


def test_synth_is_decomposition_independent():
    full = host.synth_rect(7, 300, 0, 20, 0, 300)
    part = host.synth_rect(7, 300, 5, 6, 17, 100)
    assert np.array_equal(part, full[5:11, 17:117])


REF_GREY = "/root/reference/waterfall_grey_1920_2520.raw"
REF_RGB = "/root/reference/waterfall_1920_2520.raw"


@pytest.mark.skipif(not os.path.exists(REF_GREY), reason="reference data assets not mounted")
def test_reference_assets_have_the_documented_format():
    """headerless row-major uint8, 1920 (W) x 2520 (H), 1 / 3 channels."""
    assert host.file_size(REF_GREY) == 1920 * 2520
    assert host.file_size(REF_RGB) == 1920 * 2520 * 3
    strip = host.read_rect(REF_GREY, 1920, 1000, 64, 0, 1920)
    f = Filter.builtin("gaussian")
    # a strip of the real image through both CPU implementations
    assert np.array_equal(host.cpu_run_image(strip, f.spec, 3), host.oracle(strip, f.spec, 3))


def test_native_planner_invariants():
    from parallel_convolution_b200.parallel.plan import plan_persist, plan_stream
    for rows, cb, C, kind, loops in [(2520, 5760, 3, 2, 100), (2520, 1920, 1, 2, 100), (630, 2880, 3, 1, 7),
                                     (16384, 16384, 1, 0, 500), (37, 41, 1, 2, 3)]:
        p = plan_stream(rows, cb, C, kind, loops)
        assert p.TH + 2 * p.k <= 16 * p.R and p.HB >= p.k * C and p.HB % 16 == 0
        assert p.TWB + 2 * p.HB == 512 and p.ntx * p.TWB >= cb and p.nty * p.TH >= rows
        assert sum(p.launches(loops)) == loops
    for warps in (8, 16):
        r = plan_persist(2520, 5760, 3, 2, 100, warps=warps)
        assert r is not None and r.ntx * r.nty <= 148 * (16 // warps) and r.TH >= r.k and r.W == warps
    assert plan_persist(16384, 16384, 1, 2, 100) is None      # does not fit on chip -> stream


@pytest.mark.parametrize("name", ["gaussian", "box", "edge", "sharpen", "sobel_x"])
def test_cpu_engine_matches_plain_torch_fp32(rng, name):
    """Same plain-PyTorch fp32 reference the GPU tests use (tests/test_gpu_vs_torch.py)."""
    from test_gpu_vs_torch import torch_reference
    img = random_image(rng, 33, 47, 3)
    f = Filter.builtin(name)
    ref = torch_reference(img, f.taps, 3, "cpu")
    assert np.array_equal(host.cpu_run_image(img, f.spec, 3), ref)
    assert np.array_equal(host.oracle(img, f.spec, 3), ref)


@pytest.mark.parametrize("k", [1, 3, 7])
def test_cpu_temporal_blocking_is_bit_identical(rng, k):
    img = random_image(rng, 150, 97, 3)
    for name in ("gaussian", "edge", "laplacian"):
        f = Filter.builtin(name)
        assert np.array_equal(host.cpu_run_image_blocked(img, f.spec, 10, k, threads=3),
                              host.oracle(img, f.spec, 10)), (name, k)
