#!/usr/bin/env python
"""`mpirun -np 4 ./mpi_conv image W H loops rgb` without MPI (and without a GPU).

Two launchers, same result:
  * `python -m parallel_convolution_b200 -np 4 ... --backend cpu`  — ranks under torchrun; on one node they share
    the image in /dev/shm and exchange only epoch counters (--cpu-transport shm), across nodes gloo messages;
  * `pconv_native -np 4 ... --cpu` — the C++ binary forks the ranks itself, no Python.
"""
import os
import subprocess
import sys
import tempfile

import numpy as np

from parallel_convolution_b200 import Filter
from parallel_convolution_b200.ops import host

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NATIVE = os.path.join(ROOT, "parallel_convolution_b200", "_lib", "pconv_native")


def main():
    w, h, loops = 480, 360, 12
    img = np.random.default_rng(2).integers(0, 256, (h, w, 3), dtype=np.uint8)
    want = host.oracle(img, Filter.builtin("gaussian").spec, loops).reshape(img.shape)
    with tempfile.TemporaryDirectory() as d:
        img.tofile(os.path.join(d, "in.raw"))
        env = dict(os.environ, PYTHONPATH=ROOT)
        p = subprocess.run([sys.executable, "-m", "parallel_convolution_b200", "-np", "4", "in.raw", str(w), str(h),
                            str(loops), "rgb", "--backend", "cpu", "-o", "py.raw"], cwd=d, env=env,
                           capture_output=True, text=True, timeout=300)
        assert p.returncode == 0, p.stderr
        print("python -np 4 : loop", p.stdout.strip().splitlines()[-1], "s, equals the oracle:",
              np.array_equal(np.fromfile(os.path.join(d, "py.raw"), dtype=np.uint8).reshape(img.shape), want))
        if os.path.exists(NATIVE):
            q = subprocess.run([NATIVE, "-np", "4", "in.raw", str(w), str(h), str(loops), "rgb", "--cpu", "-o", "nat.raw"],
                               cwd=d, capture_output=True, text=True, timeout=60)
            assert q.returncode == 0, q.stderr
            print("native -np 4 : loop", q.stdout.strip().splitlines()[-1], "s, equals the oracle:",
                  np.array_equal(np.fromfile(os.path.join(d, "nat.raw"), dtype=np.uint8).reshape(img.shape), want))


if __name__ == "__main__":
    main()
