"""ConvEngine public API on the CPU backend (no GPU): run(), tile I/O, file round trip, reports."""
import os

import numpy as np
import pytest

from conftest import random_image
from parallel_convolution_b200 import ConvEngine, Filter, available_filters
from parallel_convolution_b200.ops import host


@pytest.mark.parametrize("c", [1, 3])
def test_run_matches_oracle(rng, c):
    img = random_image(rng, 45, 52, c)
    with ConvEngine(width=52, height=45, channels=c, filter="gaussian", backend="cpu") as eng:
        out = eng.run(img, 6)
        assert eng.backend == "cpu" and eng.report.loops == 6 and eng.report.elapsed_ms > 0
    assert out.shape == img.shape
    assert np.array_equal(out, host.oracle(img, Filter.builtin("gaussian").spec, 6))


def test_run_file_prefixes_basename_and_handles_zero_loops(tmp_path, rng):
    img = random_image(rng, 30, 40, 3)
    src = tmp_path / "sub" / "img.raw"
    src.parent.mkdir()
    img.tofile(src)
    with ConvEngine(width=40, height=30, channels=3, filter="sharpen", backend="cpu") as eng:
        eng.run_file(os.fspath(src), 0)
    out = np.fromfile(src.parent / "blur_img.raw", dtype=np.uint8).reshape(img.shape)
    assert np.array_equal(out, img)                   # loops=0: the image itself (any parity works)


def test_synthetic_load_equals_host_generator():
    with ConvEngine(width=33, height=21, channels=3, filter="identity", backend="cpu") as eng:
        eng.load_synthetic(5)
        eng.iterate(1)
        got = eng.gather()
    assert np.array_equal(got.reshape(21, 99), host.synth_rect(5, 99, 0, 21, 0, 99))


def test_unknown_backend_and_filter_are_rejected():
    with pytest.raises(ValueError):
        ConvEngine(width=8, height=8, backend="tpu")
    with pytest.raises((KeyError, ValueError)):
        ConvEngine(width=8, height=8, filter="not-a-filter", backend="cpu")
    assert {"box", "gaussian", "edge"} <= set(available_filters())     # the reference's three tables


def test_custom_filter_objects(rng):
    img = random_image(rng, 20, 20, 1)
    f = Filter.from_int([1, 1, 1, 1, 8, 1, 1, 1, 1], 16, name="mine")
    with ConvEngine(width=20, height=20, channels=1, filter=f, backend="cpu") as eng:
        assert np.array_equal(eng.run(img, 3), host.oracle(img, f.spec, 3))


def test_usable_connections_follow_the_environment(monkeypatch):
    """The package never overrides a CUDA_DEVICE_MAX_CONNECTIONS the user chose; the runtime sizes its
    per-virtual-rank streams from what is really in effect (and falls back to streaming beyond it)."""
    from parallel_convolution_b200.ops import cuda as K
    monkeypatch.setenv("CUDA_DEVICE_MAX_CONNECTIONS", "1")
    assert K.usable_connections() == 1
    monkeypatch.setenv("CUDA_DEVICE_MAX_CONNECTIONS", "64")
    assert K.usable_connections() == K.MAX_CONNECTIONS
    monkeypatch.setenv("CUDA_DEVICE_MAX_CONNECTIONS", "bogus")
    assert K.usable_connections() == 8
    monkeypatch.delenv("CUDA_DEVICE_MAX_CONNECTIONS")
    assert K.usable_connections() == 8


def test_runtime_rejects_unknown_modes_and_warp_counts():
    from parallel_convolution_b200.engine import EngineConfig
    cfg = EngineConfig(width=64, height=48, mode="resident", warps=8)
    assert cfg.mode == "resident" and cfg.warps == 8 and not cfg.standalone
    import inspect
    from parallel_convolution_b200.parallel import fused_runtime
    src = inspect.getsource(fused_runtime.FusedRuntime.__init__)
    assert "unknown kernel mode" in src and "8 or 16" in src
