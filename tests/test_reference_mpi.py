"""The reference's CPU program (mpi/mpi_conv.c, unmodified) under the single-node MPI stand-in of
``baseline/mpi_shim``: the stand-in's self-test, and the cross-check that the reference with one process
produces exactly the bytes of this framework's oracle / CPU engine (same zero padding, same fp32 tap order,
same truncation: SURVEY.md §2.4), while its multi-process output differs near tile junctions (defects D1/D2,
SURVEY.md §2.5) — which is why the framework's own np>1 results are checked against np=1, not against
``mpirun -np N ./mpi_conv``.
"""
import os
import shutil
import subprocess

import numpy as np
import pytest

from parallel_convolution_b200.models.filters import Filter
from parallel_convolution_b200.ops import host as H

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHIM = os.path.join(ROOT, "baseline", "mpi_shim")
REF_SRC = os.path.join(os.environ.get("PCONV_REFERENCE", "/root/reference"), "mpi", "mpi_conv.c")

pytestmark = pytest.mark.skipif(shutil.which("gcc") is None, reason="needs gcc")


def _cc(args, cwd):
    p = subprocess.run(["gcc", "-O2", "-I" + SHIM] + args, cwd=cwd, capture_output=True, text=True)
    assert p.returncode == 0, p.stderr[-2000:]


@pytest.fixture(scope="module")
def shim(tmp_path_factory):
    d = tmp_path_factory.mktemp("mpi_shim")
    _cc(["-o", "mpirun_shim", os.path.join(SHIM, "mpirun_shim.c")], d)
    _cc(["-o", "selftest", os.path.join(SHIM, "selftest.c"), os.path.join(SHIM, "mpi_shim.c")], d)
    return d


@pytest.mark.parametrize("np_", [1, 2, 3, 4, 9])
def test_shim_selftest(shim, np_):
    cmd = ["./selftest"] if np_ == 1 else ["./mpirun_shim", "-np", str(np_), "./selftest"]
    p = subprocess.run(cmd, cwd=shim, capture_output=True, text=True, timeout=120)
    assert p.returncode == 0, p.stdout + p.stderr
    assert f"selftest ok {np_}" in p.stdout


def test_launcher_propagates_failure(shim):
    """A rank that aborts must stop the job with a non-zero status (MPI_Abort on bad arguments)."""
    src = shim / "abort.c"
    src.write_text('#include <mpi.h>\nint main(int c,char**v){int r;MPI_Init(&c,&v);MPI_Comm_rank(MPI_COMM_WORLD,&r);'
                   'if(r==1)MPI_Abort(MPI_COMM_WORLD,5);MPI_Barrier(MPI_COMM_WORLD);MPI_Finalize();return 0;}\n')
    _cc(["-o", "abort", "abort.c", os.path.join(SHIM, "mpi_shim.c")], shim)
    p = subprocess.run(["./mpirun_shim", "-np", "3", "./abort"], cwd=shim, capture_output=True, text=True, timeout=60)
    assert p.returncode == 5


@pytest.fixture(scope="module")
def mpi_conv(shim):
    if not os.path.exists(REF_SRC):
        pytest.skip("reference sources not present")
    shutil.copyfile(REF_SRC, shim / "mpi_conv.c")
    _cc(["-w", "-o", "mpi_conv", "mpi_conv.c", os.path.join(SHIM, "mpi_shim.c")], shim)
    return shim


def _run_reference(d, img: np.ndarray, loops: int, np_: int) -> np.ndarray:
    h, w = img.shape[:2]
    color = "rgb" if img.ndim == 3 and img.shape[2] == 3 else "grey"
    name = f"in_{w}_{h}_{color}.raw"
    img.tofile(d / name)
    out = d / ("blur_" + name)
    if out.exists():
        out.unlink()
    cmd = ["./mpirun_shim", "-np", str(np_), "./mpi_conv", name, str(w), str(h), str(loops), color]
    p = subprocess.run(cmd, cwd=d, capture_output=True, text=True, timeout=120)
    assert p.returncode == 0, p.stdout + p.stderr
    float(p.stdout.strip().splitlines()[-1])            # it prints its loop time
    return np.fromfile(out, dtype=np.uint8).reshape(img.shape)


@pytest.mark.parametrize("shape,loops", [((32, 48), 5), ((40, 36, 3), 7)])
def test_reference_single_process_equals_oracle(mpi_conv, shape, loops):
    img = np.random.default_rng(3).integers(0, 256, shape, dtype=np.uint8)
    ref = _run_reference(mpi_conv, img, loops, 1)
    spec = Filter.builtin("gaussian").spec
    assert np.array_equal(ref, H.oracle(img, spec, loops).reshape(img.shape))
    assert np.array_equal(ref, H.cpu_run_image(img, spec, loops).reshape(img.shape))


def test_reference_multi_process_differs_only_near_junctions(mpi_conv):
    """2x2 ranks: the reference never exchanges the diagonal halo cells (D1), so after a few iterations its
    output differs from the single-process result around the interior junction and nowhere far from tile
    edges; one process column or row apart the results agree."""
    img = np.random.default_rng(4).integers(0, 256, (32, 48), dtype=np.uint8)
    loops = 3
    want = H.oracle(img, Filter.builtin("gaussian").spec, loops).reshape(img.shape)
    got = _run_reference(mpi_conv, img, loops, 4)
    ys, xs = np.nonzero(got != want)
    assert len(ys) > 0                                   # the defect is real
    d_edge = np.minimum(np.abs(ys - 15.5), np.abs(xs - 23.5))
    assert d_edge.max() <= loops + 0.5                   # and confined to the tile cuts
