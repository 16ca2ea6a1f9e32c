"""The programs under examples/ run (on the CPU backend here) and check themselves against the oracle."""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(name, *args, cwd=None):
    env = dict(os.environ, PYTHONPATH=ROOT, PCONV_ALLOW_CPU_FALLBACK="1", CUDA_VISIBLE_DEVICES="")
    return subprocess.run([sys.executable, os.path.join(ROOT, "examples", name), *map(str, args)], cwd=cwd or ROOT,
                          env=env, capture_output=True, text=True, timeout=300)


@pytest.mark.parametrize("name", ["custom_filter.py", "decomposition_invariance.py", "cpu_ranks.py"])
def test_example_runs_and_verifies_itself(name):
    p = _run(name)
    assert p.returncode == 0, p.stdout[-1500:] + p.stderr[-1500:]
    assert "False" not in p.stdout


def test_blur_file_example_matches_the_cli_contract(tmp_path):
    img = np.random.default_rng(0).integers(0, 256, (40, 56), dtype=np.uint8)
    img.tofile(tmp_path / "pic.raw")
    p = _run("blur_file.py", "pic.raw", 56, 40, 3, "grey", cwd=tmp_path)
    assert p.returncode == 0, p.stderr
    float(p.stdout.strip())
    from parallel_convolution_b200 import Filter
    from parallel_convolution_b200.ops import host
    got = np.fromfile(tmp_path / "blur_pic.raw", dtype=np.uint8).reshape(40, 56)
    assert np.array_equal(got, host.oracle(img, Filter.builtin("gaussian").spec, 3).reshape(40, 56))
