"""The command-line programs on a real GPU: Python CLI (fused backend, --verify against the CPU
engine), the stand-alone native binary, and parity between the two."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _cli(*args):
    env = dict(os.environ, PYTHONPATH=ROOT)
    return subprocess.run([sys.executable, "-m", "parallel_convolution_b200", *map(str, args)],
                          capture_output=True, text=True, cwd=ROOT, env=env, timeout=300)


@pytest.mark.parametrize("color,filt", [("rgb", "gaussian"), ("grey", "sharpen"), ("rgb", "box")])
def test_cli_fused_verify(tmp_path, color, filt):
    r = _cli("syn.raw", 640, 400, 25, color, "--backend", "fused", "--synthetic", 4, "--filter", filt,
             "--verify", "--json", "-o", tmp_path / "o.raw")
    assert r.returncode == 0, r.stderr + r.stdout
    rep = json.loads(r.stdout)
    assert rep["verified"] is True and rep["backend"] == "fused" and rep["kernel_launches"] > 0


def test_cli_virtual_ranks_and_autotune(tmp_path):
    r = _cli("syn.raw", 640, 400, 25, "rgb", "--backend", "fused", "--synthetic", 4, "--workers", 4,
             "--autotune", "--verify", "--json", "--no-output")
    assert r.returncode == 0, r.stderr + r.stdout
    assert json.loads(r.stdout)["verified"] is True


def test_native_binary_matches_python_cli(tmp_path):
    exe = os.path.join(ROOT, "parallel_convolution_b200", "_lib", "pconv_native")
    if not os.path.exists(exe):
        pytest.skip("pconv_native not built")
    rng = np.random.default_rng(9)
    img = rng.integers(0, 256, (300, 500, 3), dtype=np.uint8)
    src = tmp_path / "in.raw"
    img.tofile(src)
    n = subprocess.run([exe, str(src), "500", "300", "13", "rgb", "-o", str(tmp_path / "n.raw")],
                       capture_output=True, text=True, timeout=300)
    assert n.returncode == 0 and "Execution time" in n.stdout, n.stderr + n.stdout
    p = _cli(src, 500, 300, 13, "rgb", "--backend", "fused", "-o", tmp_path / "p.raw")
    assert p.returncode == 0, p.stderr
    assert np.array_equal(np.fromfile(tmp_path / "n.raw", dtype=np.uint8),
                          np.fromfile(tmp_path / "p.raw", dtype=np.uint8))
