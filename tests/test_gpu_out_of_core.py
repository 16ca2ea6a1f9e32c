"""GPU out-of-core mode (utils/out_of_core.py): an image "larger than device memory" (the limit is forced
small) streams through the GPU in row strips with k-deep halos, file -> file, bit-identical to the CPU engine."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import random_image
from parallel_convolution_b200.models.filters import Filter
from parallel_convolution_b200.ops import host

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("name,loops,k,mem_kib", [("gaussian", 21, 8, 700), ("sharpen", 9, 4, 300), ("box", 5, 5, 2000)])
def test_strips_match_cpu_engine(rng, tmp_path, name, loops, k, mem_kib):
    from parallel_convolution_b200.utils import out_of_core as ooc
    h, w, c = 700, 333, 3
    img = random_image(rng, h, w, c)
    src, dst = tmp_path / "in.raw", tmp_path / "out.raw"
    img.tofile(src)
    f = Filter.builtin(name)
    info = ooc.run_file(str(src), str(dst), w, h, c, f, loops, mem_kib << 10, k=k)
    assert info["strips"] >= 2 and info["passes"] == -(-loops // k)
    out = np.fromfile(dst, dtype=np.uint8).reshape(img.shape)
    assert np.array_equal(out, host.cpu_run_image(img, f.spec, loops))
    assert not os.path.exists(str(dst) + ".pass")


def test_cli_out_of_core_on_gpu(rng, tmp_path):
    h, w = 420, 512
    img = random_image(rng, h, w, 1)
    src = tmp_path / "g.raw"
    img.tofile(src)
    r = subprocess.run([sys.executable, "-m", "parallel_convolution_b200", str(src), str(w), str(h), "13", "grey",
                        "--backend", "fused", "--out-of-core", "1", "--halo-depth", "5", "--json",
                        "-o", str(tmp_path / "o.raw")], capture_output=True, text=True, cwd=ROOT,
                       env=dict(os.environ, PYTHONPATH=ROOT), timeout=300)
    assert r.returncode == 0, r.stderr
    rep = json.loads(r.stdout)
    assert rep["backend"].startswith("fused") and rep["strips"] >= 2 and rep["passes"] == 3
    out = np.fromfile(tmp_path / "o.raw", dtype=np.uint8).reshape(h, w)
    assert np.array_equal(out, host.cpu_run_image(img, Filter.builtin("gaussian").spec, 13))
