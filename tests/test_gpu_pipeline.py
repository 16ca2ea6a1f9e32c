"""Pipelined end-to-end steps (H2D || iterate || D2H on rotating buffer slots): every step's result
must equal the oracle of ITS input, on one tile and on virtual ranks, in both kernel modes."""
import numpy as np
import pytest

from conftest import random_image
from parallel_convolution_b200.models.filters import Filter
from parallel_convolution_b200.ops import host

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("workers,mode", [(1, "stream"), (1, "resident"), (4, "stream"), (4, "resident")])
def test_pipelined_steps_match_oracle(rng, workers, mode):
    from parallel_convolution_b200 import ConvEngine
    H, W, C, loops, slots, steps = 240, 320, 3, 7, 3, 8
    f = Filter.builtin("gaussian")
    with ConvEngine(width=W, height=H, channels=C, filter=f, backend="fused", workers=workers,
                    mode=mode, slots=slots, halo_depth=3) as eng:
        stage = [eng.pinned_buffers() for _ in range(slots)]
        images = [random_image(rng, H, W, C) for _ in range(steps)]
        results = []
        for i, img in enumerate(images):
            ins, outs = stage[i % slots]
            if i >= slots:               # the slot's previous result must be consumed before reuse
                eng.synchronize()
                results.append({r: o.copy() for r, o in stage[i % slots][1].items()})
            flat = img.reshape(H, W * C)
            for r, arr in ins.items():
                t = eng.grid.tile(r)
                arr[:] = flat[t.row0:t.row0 + t.rows, t.col0 * C:(t.col0 + t.cols) * C]
            eng.enqueue_step_pipelined(ins, outs, loops)
        eng.synchronize()
        done = len(results)
        for i in range(done, steps):
            results.append({r: o.copy() for r, o in stage[i % slots][1].items()})
        for i, img in enumerate(images):
            ref = host.cpu_run_image(img, f.spec, loops).reshape(H, W * C)
            for r, tile in results[i].items():
                t = eng.grid.tile(r)
                assert np.array_equal(tile, ref[t.row0:t.row0 + t.rows, t.col0 * C:(t.col0 + t.cols) * C]), \
                    (workers, mode, i, r)


@pytest.mark.parametrize("workers", [1, 4])
def test_autotune_keeps_results_exact(rng, workers):
    from parallel_convolution_b200 import ConvEngine
    img = random_image(rng, 360, 480, 3)
    f = Filter.builtin("gaussian")
    with ConvEngine(width=480, height=360, channels=3, filter=f, backend="fused", workers=workers) as eng:
        info = eng.autotune(20)
        assert info and info["mode"] in ("stream", "resident") and len(info["tried"]) >= 2
        out = eng.run(img, 20)
    assert np.array_equal(out, host.cpu_run_image(img, f.spec, 20))


def test_batch_mode_streams_files(rng, tmp_path):
    from parallel_convolution_b200 import ConvEngine
    H, W, C, loops = 120, 160, 3, 6
    f = Filter.builtin("sharpen")
    paths, imgs = [], []
    for i in range(5):
        img = random_image(rng, H, W, C)
        p = tmp_path / f"b{i}.raw"
        img.tofile(p)
        paths.append(str(p))
        imgs.append(img)
    with ConvEngine(width=W, height=H, channels=C, filter=f, backend="fused", workers=2, slots=3) as eng:
        info = eng.run_batch(paths, loops, str(tmp_path))
    assert info["pipelined"] and info["images"] == 5
    for p, img in zip(paths, imgs):
        got = np.fromfile(tmp_path / ("blur_" + p.split("/")[-1]), dtype=np.uint8).reshape(img.shape)
        assert np.array_equal(got, host.cpu_run_image(img, f.spec, loops))
