"""CLI parity with the reference programs (argument grammar, blur_ output, printed seconds) and the
extras: --json, --synthetic, --verify, checkpoint/resume, error handling. CPU only."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

from parallel_convolution_b200.models.filters import Filter
from parallel_convolution_b200.ops import host

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run_cli(*args, cwd=None):
    env = dict(os.environ, PYTHONPATH=ROOT)
    return subprocess.run([sys.executable, "-m", "parallel_convolution_b200", *map(str, args)],
                          capture_output=True, text=True, cwd=cwd or ROOT, env=env, timeout=180)


@pytest.fixture()
def raw(tmp_path):
    rng = np.random.default_rng(11)
    img = rng.integers(0, 256, (40, 56), dtype=np.uint8)
    p = tmp_path / "pic_56_40.raw"
    img.tofile(p)
    return p, img


def test_reference_grammar_and_blur_prefix(raw):
    p, img = raw
    r = run_cli(p, 56, 40, 7, "grey", "--backend", "cpu")
    assert r.returncode == 0, r.stderr
    assert float(r.stdout.strip()) >= 0.0
    out = p.parent / ("blur_" + p.name)          # prefix on the BASENAME (reference defect D8 fixed)
    got = np.fromfile(out, dtype=np.uint8).reshape(40, 56)
    assert np.array_equal(got, host.oracle(img, Filter.builtin("gaussian").spec, 7))


def test_usage_error_exit_code(raw):
    p, _ = raw
    r = run_cli(p, 56, 40, 7)                    # missing rgb/grey
    assert r.returncode != 0 and "image_name width height loops [rgb/grey]" in r.stderr
    r = run_cli(p, 56, 40, 7, "purple")
    assert r.returncode != 0


def test_short_file_fails_fast(raw):
    p, _ = raw
    r = run_cli(p, 560, 400, 1, "grey", "--backend", "cpu")
    assert r.returncode == 1 and "need" in r.stderr


def test_json_synthetic_verify_filter(tmp_path):
    r = run_cli("synthetic.raw", 37, 29, 4, "rgb", "--backend", "cpu", "--synthetic", 3, "--json",
                "--filter", "0,-1,0,-1,5,-1,0,-1,0", "--verify", "-o", tmp_path / "o.raw")
    assert r.returncode == 0, r.stderr
    rep = json.loads(r.stdout)
    assert rep["verified"] is True and rep["filter"]["path"] == "int-swar" and rep["iterations"] == 4
    assert os.path.getsize(tmp_path / "o.raw") == 37 * 29 * 3


def test_checkpoint_and_resume_are_bit_identical(raw, tmp_path):
    p, img = raw
    full = tmp_path / "full.raw"
    assert run_cli(p, 56, 40, 10, "grey", "--backend", "cpu", "-o", full).returncode == 0
    part = tmp_path / "part.raw"
    # run 10 iterations with a checkpoint every 4, then pretend the job died after the second one
    r = run_cli(p, 56, 40, 10, "grey", "--backend", "cpu", "-o", part, "--checkpoint-every", 4, "--no-output")
    assert r.returncode == 0, r.stderr
    side = json.loads(open(str(part) + ".ckpt.json").read())
    assert side["iterations_done"] == 8
    r = run_cli(p, 56, 40, 10, "grey", "--backend", "cpu", "-o", part, "--resume")
    assert r.returncode == 0, r.stderr
    assert np.array_equal(np.fromfile(part, dtype=np.uint8), np.fromfile(full, dtype=np.uint8))
    assert not os.path.exists(str(part) + ".ckpt.json")    # cleaned up after a complete run


def test_checkpoint_is_crash_safe_after_the_first_one(raw, tmp_path):
    """The second checkpoint goes to a new file; the sidecar always names a complete image whose size and
    CRC are recorded — a truncated / corrupted state image is refused instead of silently resumed."""
    p, img = raw
    part = tmp_path / "part.raw"
    r = run_cli(p, 56, 40, 10, "grey", "--backend", "cpu", "-o", part, "--checkpoint-every", 4, "--no-output")
    assert r.returncode == 0, r.stderr
    side = json.loads(open(str(part) + ".ckpt.json").read())
    state_img = tmp_path / side["image"]
    assert side["iterations_done"] == 8 and side["image"].endswith(".ckpt.8.raw")
    assert side["bytes"] == 56 * 40 and os.path.getsize(state_img) == 56 * 40
    assert not os.path.exists(str(part) + ".ckpt.4.raw")          # the previous image went away last
    # a crash while the NEXT checkpoint image is being written leaves this pair intact: simulate the
    # half-written newer file and make sure the resume still uses the complete one
    (tmp_path / "part.raw.ckpt.12.raw").write_bytes(b"\x00" * 100)
    full = tmp_path / "full.raw"
    assert run_cli(p, 56, 40, 10, "grey", "--backend", "cpu", "-o", full).returncode == 0
    r = run_cli(p, 56, 40, 10, "grey", "--backend", "cpu", "-o", part, "--resume")
    assert r.returncode == 0, r.stderr
    assert np.array_equal(np.fromfile(part, dtype=np.uint8), np.fromfile(full, dtype=np.uint8))
    # a state image that does not match its sidecar (bit rot / interrupted overwrite) is refused
    r = run_cli(p, 56, 40, 10, "grey", "--backend", "cpu", "-o", part, "--checkpoint-every", 4, "--no-output")
    assert r.returncode == 0, r.stderr
    side = json.loads(open(str(part) + ".ckpt.json").read())
    data = bytearray((tmp_path / side["image"]).read_bytes())
    data[7] ^= 0xFF
    (tmp_path / side["image"]).write_bytes(bytes(data))
    r = run_cli(p, 56, 40, 10, "grey", "--backend", "cpu", "-o", part, "--resume")
    assert r.returncode != 0 and "checksum" in (r.stderr + r.stdout)


def test_batch_mode_cpu(tmp_path):
    rng = np.random.default_rng(2)
    paths = []
    for i in range(3):
        img = rng.integers(0, 256, (24, 32, 3), dtype=np.uint8)
        p = tmp_path / f"im{i}.raw"
        img.tofile(p)
        paths.append((p, img))
    lst = tmp_path / "list.txt"
    lst.write_text("\n".join(str(p) for p, _ in paths) + "\n")
    outd = tmp_path / "out"
    outd.mkdir()
    r = run_cli("ignored", 32, 24, 4, "rgb", "--backend", "cpu", "--batch", lst, "-o", outd, "--json")
    assert r.returncode == 0, r.stderr
    assert json.loads(r.stdout)["images"] == 3
    f = Filter.builtin("gaussian")
    for p, img in paths:
        got = np.fromfile(outd / ("blur_" + p.name), dtype=np.uint8).reshape(img.shape)
        assert np.array_equal(got, host.oracle(img, f.spec, 4))


def test_np_flag_spawns_ranks_like_mpirun(raw):
    """`pconv -np 2 ...` == `mpirun -np 2 ./mpi_conv ...`: two gloo ranks, one output, oracle bytes."""
    p, img = raw
    r = run_cli("-np", 2, p, 56, 40, 5, "grey", "--backend", "cpu", "--json")
    assert r.returncode == 0, r.stderr
    rep = json.loads(r.stdout.strip().splitlines()[-1])
    assert rep["workers"] == 2
    got = np.fromfile(p.parent / ("blur_" + p.name), dtype=np.uint8).reshape(40, 56)
    assert np.array_equal(got, host.oracle(img, Filter.builtin("gaussian").spec, 5))


def test_out_of_core_file_run_matches_oracle(tmp_path):
    """--out-of-core: strips + k-deep halos + several passes over ping-pong files == the oracle."""
    rng = np.random.default_rng(23)
    img = rng.integers(0, 256, (300, 67, 3), dtype=np.uint8)
    src = tmp_path / "big.raw"
    img.tofile(src)
    ref = host.oracle(img, Filter.builtin("sharpen").spec, 11)
    # direct binding: tiny memory budget -> many strips; k=4 -> 3 passes (4+4+3)
    out = tmp_path / "o1.raw"
    host.cpu_run_file(str(src), str(out), 300, 67, 3, Filter.builtin("sharpen").spec, 11, k=4, mem_bytes=40_000)
    assert np.array_equal(np.fromfile(out, dtype=np.uint8).reshape(img.shape), ref)
    assert not os.path.exists(str(out) + ".pass")
    # CLI, default k, even number of passes, grey
    g = rng.integers(0, 256, (257, 130), dtype=np.uint8)
    gs = tmp_path / "g.raw"
    g.tofile(gs)
    r = run_cli(gs, 130, 257, 16, "grey", "--backend", "cpu", "--out-of-core", 1, "--json", "-o", tmp_path / "o2.raw")
    assert r.returncode == 0, r.stderr
    assert json.loads(r.stdout)["backend"].startswith("cpu")
    got = np.fromfile(tmp_path / "o2.raw", dtype=np.uint8).reshape(g.shape)
    assert np.array_equal(got, host.oracle(g, Filter.builtin("gaussian").spec, 16))
    # loops = 0 copies; in == out is refused
    host.cpu_run_file(str(gs), str(tmp_path / "o3.raw"), 257, 130, 1, Filter.builtin("box").spec, 0)
    assert np.array_equal(np.fromfile(tmp_path / "o3.raw", dtype=np.uint8).reshape(g.shape), g)
    with pytest.raises(OSError):
        host.cpu_run_file(str(gs), str(gs), 257, 130, 1, Filter.builtin("box").spec, 3)


def test_netpbm_round_trip(tmp_path):
    from parallel_convolution_b200.utils import netpbm
    rng = np.random.default_rng(2)
    for c in (1, 3):
        img = rng.integers(0, 256, (17, 23, c) if c == 3 else (17, 23), dtype=np.uint8)
        raw, pnm, back = tmp_path / f"a{c}.raw", tmp_path / f"a{c}.pnm", tmp_path / f"b{c}.raw"
        img.tofile(raw)
        assert netpbm.main(["raw2pnm", str(raw), "23", "17", "rgb" if c == 3 else "grey", str(pnm)]) == 0
        head = open(pnm, "rb").read(15)
        assert head.startswith(b"P6\n23 17\n255\n" if c == 3 else b"P5\n23 17\n255\n")
        assert netpbm.pnm_to_raw(str(pnm), str(back)) == (23, 17, c)
        assert np.array_equal(np.fromfile(back, dtype=np.uint8), img.ravel())
    assert netpbm.main(["raw2pnm", str(raw), "230", "170", "rgb", str(pnm)]) == 1   # short input
