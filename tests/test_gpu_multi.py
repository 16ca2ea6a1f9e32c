"""Real multi-GPU tests (one process per GPU over NVLink peer memory). Skipped on boxes with fewer
than two GPUs — the same logic is covered on one GPU by the virtual-rank tests."""
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpus() -> int:
    from parallel_convolution_b200.ops import cuda as K
    return K.device_count()


def _port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _run(nproc, args, **extra_env):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}",
           "--master-addr", "127.0.0.1", "--master-port", str(_port()),
           os.path.join(ROOT, "scripts", "dist_check.py")] + args
    env = dict(os.environ, PYTHONPATH=ROOT, **extra_env)
    return subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT, env=env)


@pytest.mark.parametrize("backend", ["fused", "nccl"])
@pytest.mark.parametrize("extra", [[], ["--filter", "sharpen", "--channels", "1", "--halo-depth", "3"],
                                   ["--filter", "box", "--height", "250", "--width", "333"]])
def test_multi_gpu_matches_cpu(backend, extra):
    n = _ngpus()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    nproc = 8 if n >= 8 else (4 if n >= 4 else 2)
    r = _run(nproc, ["--backend", backend] + extra)
    assert "DIST_CHECK OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]


def test_nccl_backend_with_poisoned_ghosts():
    """PCONV_POISON_HALO=1: ghost strips are refilled with 0xA5 before every exchange, so anything the
    kernel consumes without it having been delivered by this exchange would corrupt the result."""
    n = _ngpus()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    nproc = 8 if n >= 8 else (4 if n >= 4 else 2)
    r = _run(nproc, ["--backend", "nccl", "--halo-depth", "2"], PCONV_POISON_HALO="1")
    assert "DIST_CHECK OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
