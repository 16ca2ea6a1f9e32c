"""Cross-check against the UNMODIFIED reference MPI program (CPU only, no GPU needed).

Runs baseline/_ref/mpi/mpi_conv_O3 (mpi/mpi_conv.c verbatim, linked against the single-node MPI stand-in
baseline/mpi_shim) on the reference's own test images with 1, 2, 4 and 8 processes and compares every byte
with this framework's CPU backend (shared-image ranks, same process counts) and with the oracle.

Expected, and checked:
  * reference with ONE process == oracle == every one of our runs, byte for byte (same zero padding, fp32 tap
    order and truncation);
  * reference with SEVERAL processes differs from its own single-process result near the tile cuts — the
    diagonal halo cells it never exchanges (defect D1) and the border corners it never recomputes (D2),
    SURVEY.md §2.5 — while our np > 1 runs stay identical to np = 1.
"""
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from parallel_convolution_b200.models.filters import Filter  # noqa: E402
from parallel_convolution_b200.ops import host as H  # noqa: E402

MPI = os.path.join(ROOT, "baseline", "_ref", "mpi")
NATIVE = os.path.join(ROOT, "parallel_convolution_b200", "_lib", "pconv_native")


def main() -> int:
    if not os.path.exists(os.path.join(MPI, "mpi_conv_O3")):
        subprocess.run([sys.executable, os.path.join(ROOT, "baseline", "build_reference.py")], check=True)
    work = "/tmp/pconv_crosscheck_mpi"
    os.makedirs(work, exist_ok=True)
    W, Hh, L = 1920, 2520, 20
    spec = Filter.builtin("gaussian").spec
    ok = True
    for color, C, name in (("grey", 1, "waterfall_grey_1920_2520.raw"), ("rgb", 3, "waterfall_1920_2520.raw")):
        src = os.path.join("/root/reference", name)
        shape = (Hh, W) if C == 1 else (Hh, W, C)
        img = (np.fromfile(src, dtype=np.uint8).reshape(shape) if os.path.exists(src)
               else np.random.default_rng(3).integers(0, 256, shape, dtype=np.uint8))
        img.tofile(os.path.join(work, name))
        want = H.oracle(img, spec, L).reshape(shape)
        for n in (1, 2, 4, 8):
            p = subprocess.run([os.path.join(MPI, "mpirun_shim"), "-np", str(n), os.path.join(MPI, "mpi_conv_O3"), name,
                                str(W), str(Hh), str(L), color], cwd=work, capture_output=True, text=True)
            if p.returncode != 0:
                print(f"{color} np={n}: reference failed: {p.stderr[-200:]}")
                ok = False
                continue
            ref = np.fromfile(os.path.join(work, "blur_" + name), dtype=np.uint8).reshape(shape)
            q = subprocess.run([NATIVE, "-np", str(n), name, str(W), str(Hh), str(L), color, "--cpu", "-o", "ours.raw"],
                               cwd=work, capture_output=True, text=True)
            ours = np.fromfile(os.path.join(work, "ours.raw"), dtype=np.uint8).reshape(shape) if q.returncode == 0 else None
            d_ref = ref != want
            px = d_ref.reshape(Hh, W, -1).any(axis=2)
            ours_bad = -1 if ours is None else int((ours != want).sum())
            print(f"{color} {W}x{Hh}, {L} iterations, {n} process(es): reference loop {p.stdout.strip()} s, "
                  f"{int(d_ref.sum())} bytes ({int(px.sum())} pixels) differ from the oracle; "
                  f"ours loop {q.stdout.strip().splitlines()[-1] if q.returncode == 0 else 'FAILED'} s, {ours_bad} bytes differ")
            ok = ok and ours_bad == 0 and (int(d_ref.sum()) == 0 if n == 1 else True)
    print("CROSSCHECK_MPI", "OK" if ok else "FAILED")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
