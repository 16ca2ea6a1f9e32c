"""Multi-process CPU tests (gloo, world_size > 1, no GPU): the host backend's 8-neighbour halo
exchange — as gloo messages and through the shared image of one node — and the CLI under torchrun. These are the tests the reference's own np>1 runs would fail
(diagonal cells never exchanged, corner pixels never recomputed: defects D1/D2)."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _torchrun(nproc: int, args: list[str], timeout: int = 240, **extra_env) -> subprocess.CompletedProcess:
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}",
           "--master-addr", "127.0.0.1", "--master-port", str(_port())] + args
    env = dict(os.environ, OMP_NUM_THREADS="1", PYTHONPATH=ROOT, **extra_env)
    return subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, cwd=ROOT, env=env)


@pytest.mark.parametrize("transport", ["gloo", "shm", "tcp"])
@pytest.mark.parametrize("nproc,extra", [
    (2, []),
    (4, ["--halo-depth", "3", "--filter", "sharpen"]),          # 2x2: exercises the diagonal cells
    (6, ["--filter", "box", "--height", "90", "--width", "131"]),  # remainder tiles
])
def test_host_backend_matches_single_process(nproc, extra, transport):
    r = _torchrun(nproc, [os.path.join(ROOT, "scripts", "dist_check.py"), "--backend", "cpu",
                          "--height", "96", "--width", "120"] + extra, PCONV_CPU_TRANSPORT=transport)
    assert "DIST_CHECK OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


@pytest.mark.parametrize("depth", [1, 3])
def test_poisoned_halos_are_always_refreshed(depth):
    """PCONV_POISON_HALO=1 overwrites every neighbour-facing halo with 0xA5 before each exchange: the
    result stays exact only if all eight strips (diagonals included) are freshly received each time."""
    r = _torchrun(4, [os.path.join(ROOT, "scripts", "dist_check.py"), "--backend", "cpu", "--height", "66",
                      "--width", "70", "--loops", "7", "--halo-depth", str(depth)], PCONV_POISON_HALO="1",
                  PCONV_CPU_TRANSPORT="gloo")
    assert "DIST_CHECK OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


def test_cli_under_torchrun_writes_same_file_as_single_process(tmp_path):
    rng = np.random.default_rng(5)
    img = rng.integers(0, 256, (64, 48, 3), dtype=np.uint8)
    src = tmp_path / "in.raw"
    img.tofile(src)
    one = tmp_path / "one.raw"
    r1 = subprocess.run([sys.executable, "-m", "parallel_convolution_b200", str(src), "48", "64", "5", "rgb",
                         "--backend", "cpu", "-o", str(one)], capture_output=True, text=True, cwd=ROOT)
    assert r1.returncode == 0, r1.stderr
    float(r1.stdout.strip())                      # the reference prints just the seconds
    four = tmp_path / "four.raw"
    r4 = _torchrun(4, ["-m", "parallel_convolution_b200", str(src), "48", "64", "5", "rgb",
                       "--backend", "cpu", "-o", str(four)])
    assert r4.returncode == 0, r4.stdout[-1500:] + r4.stderr[-1500:]
    assert np.array_equal(np.fromfile(one, dtype=np.uint8), np.fromfile(four, dtype=np.uint8))
