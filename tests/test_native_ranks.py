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


# ---- private tiles + TCP messages (csrc/host/net_ranks.cpp): the form that also spans nodes -----------------

@pytest.mark.parametrize("shape,color,np_,filt,loops,extra", [
    ((64, 96), "grey", 4, "gaussian", 7, ()),                              # 2x2: diagonal strips
    ((61, 97, 3), "rgb", 6, "sharpen", 9, ("--halo-depth", "2")),          # remainder tiles, last epoch shorter
    ((90, 40), "grey", 3, "box", 5, ("--threads", "2", "--halo-depth", "5")),   # k == loops, exact-fp32 filter
    ((48, 64, 3), "rgb", 8, "gaussian", 13, ("--halo-depth", "1")),        # the reference's schedule: k = 1
    ((30, 30), "grey", 9, "edge", 6, ("--halo-depth", "4")),               # 3x3 grid, a rank with all 8 neighbours
])
def test_tcp_ranks_match_the_oracle(tmp_path, shape, color, np_, filt, loops, extra):
    img = np.random.default_rng(sum(shape) * 3 + np_).integers(0, 256, shape, dtype=np.uint8)
    p, got = _run(tmp_path, img, loops, np_, color, ("--transport", "tcp", "--filter", filt, *extra))
    assert p.returncode == 0, p.stdout + p.stderr
    float(p.stdout.strip().splitlines()[-1])
    assert np.array_equal(got, H.oracle(img, Filter.builtin(filt).spec, loops).reshape(img.shape))


def test_tcp_ranks_started_separately_like_on_a_cluster(tmp_path):
    """One process per rank with --rank/--world/--master (what a scheduler or ssh loop does on several nodes)."""
    import socket
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    img = np.random.default_rng(11).integers(0, 256, (72, 60, 3), dtype=np.uint8)
    src, out = tmp_path / "in.raw", tmp_path / "out.raw"
    img.tofile(src)
    procs = [subprocess.Popen([EXE, str(src), "60", "72", "10", "rgb", "--cpu", "--rank", str(r), "--world", "4",
                               "--master", f"127.0.0.1:{port}", "--halo-depth", "3", "-o", str(out)],
                              stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True) for r in (3, 1, 2, 0)]
    outs = [p.communicate(timeout=120) for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    assert sum(1 for o, _ in outs if o.strip()) == 1          # only rank 0 prints the time
    got = np.fromfile(out, dtype=np.uint8).reshape(img.shape)
    assert np.array_equal(got, H.oracle(img, Filter.builtin("gaussian").spec, 10).reshape(img.shape))


def test_tcp_rank_fails_when_a_peer_never_shows_up(tmp_path):
    """A missing rank must end the job with an error, not hang it (rendezvous deadline)."""
    img = np.zeros((32, 32), dtype=np.uint8)
    src = tmp_path / "in.raw"
    img.tofile(src)
    env = dict(os.environ)
    p = subprocess.run([EXE, str(src), "32", "32", "2", "grey", "--cpu", "--rank", "1", "--world", "2",
                        "--master", "127.0.0.1:1", "--net-timeout", "2", "-o", str(tmp_path / "o.raw")],
                       capture_output=True, text=True, timeout=60, env=env)
    assert p.returncode != 0 and "cannot reach rank 0" in p.stderr
