"""`pconv_native -np N ... --cpu`: the torch-free, MPI-free counterpart of `mpirun -np N ./mpi_conv ...`
(mpi/READ_ME.txt:4-5) — N forked processes over one shared image (csrc/host/shm_ranks.cpp), rank-parallel
file reads and writes. Output must equal the oracle for any rank count, filter and image shape."""
import os
import subprocess

import numpy as np
import pytest

from parallel_convolution_b200.models.filters import Filter
from parallel_convolution_b200.ops import host as H

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "parallel_convolution_b200", "_lib", "pconv_native")

pytestmark = pytest.mark.skipif(not os.path.exists(EXE), reason="pconv_native not built (python -m parallel_convolution_b200.build)")


def _run(tmp_path, img, loops, np_, color, extra=()):
    h, w = img.shape[:2]
    src = tmp_path / f"in_{np_}.raw"
    img.tofile(src)
    out = tmp_path / f"out_{np_}.raw"
    p = subprocess.run([EXE, "-np", str(np_), str(src), str(w), str(h), str(loops), color, "--cpu", "-o", str(out),
                        *extra], capture_output=True, text=True, timeout=120)
    return p, (np.fromfile(out, dtype=np.uint8).reshape(img.shape) if out.exists() else None)


@pytest.mark.parametrize("shape,color,np_,filt,loops,extra", [
    ((64, 96), "grey", 4, "gaussian", 7, ()),                              # 2x2, default depth
    ((61, 97, 3), "rgb", 6, "sharpen", 9, ("--halo-depth", "2")),          # remainder tiles
    ((90, 40), "grey", 3, "box", 5, ("--threads", "2")),                   # exact-fp32 filter, threads in a rank
    ((48, 64, 3), "rgb", 8, "gaussian", 13, ("--halo-depth", "1")),        # one exchange per iteration
])
def test_native_ranks_match_the_oracle(tmp_path, shape, color, np_, filt, loops, extra):
    img = np.random.default_rng(sum(shape) + np_).integers(0, 256, shape, dtype=np.uint8)
    p, got = _run(tmp_path, img, loops, np_, color, ("--filter", filt, *extra))
    assert p.returncode == 0, p.stdout + p.stderr
    float(p.stdout.strip().splitlines()[-1])                  # like mpi_conv: the loop seconds, max over ranks
    assert np.array_equal(got, H.oracle(img, Filter.builtin(filt).spec, loops).reshape(img.shape))


def test_native_ranks_fail_cleanly_on_a_short_file(tmp_path):
    img = np.zeros((32, 32), dtype=np.uint8)
    src = tmp_path / "short.raw"
    img[:16].tofile(src)
    p = subprocess.run([EXE, "-np", "4", str(src), "32", "32", "3", "grey", "--cpu", "-o", str(tmp_path / "o.raw")],
                       capture_output=True, text=True, timeout=60)
    assert p.returncode != 0
    assert "shorter" in p.stderr or "cannot read" in p.stderr
