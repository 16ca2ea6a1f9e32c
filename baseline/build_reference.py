#!/usr/bin/env python
"""Build the UNMODIFIED reference into baseline/_ref/ (git-ignored; travels with gpurun).

The reference is not a Python package (no setup.py / pyproject.toml), so the prescribed
``pip install --no-index ... --target baseline/_ref /root/reference`` cannot work; the outcome of
that attempt is recorded in DESIGN.md. What *can* be built offline is its CUDA program: its sources
are copied verbatim from /root/reference/cuda and compiled with its own Makefile (no -arch: nvcc
12.9 embeds sm_52 SASS + compute_52 PTX, which the driver JIT-compiles on a B200). The MPI variants
need mpicc/mpi.h, which this image does not have.
"""
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("PCONV_REFERENCE", "/root/reference")
DST = os.path.join(ROOT, "baseline", "_ref")


def main() -> int:
    src = os.path.join(REF, "cuda")
    if not os.path.isdir(src):
        print(f"reference sources not found at {src}", file=sys.stderr)
        return 1
    out = os.path.join(DST, "cuda")
    def writable(path):
        for base, dirs, files in os.walk(path):
            for n in dirs + files:
                os.chmod(os.path.join(base, n), 0o755 if n in dirs else 0o644)
        os.chmod(path, 0o755)

    if os.path.isdir(out):
        writable(out)
        shutil.rmtree(out)
    os.makedirs(DST, exist_ok=True)
    shutil.copytree(src, out)
    writable(out)      # /root/reference is read-only; the copy must be buildable (contents untouched)
    env = dict(os.environ)
    env["PATH"] = "/usr/local/cuda/bin:" + env.get("PATH", "")
    proc = subprocess.run(["make", "-C", out], env=env, capture_output=True, text=True)
    with open(os.path.join(DST, "build.log"), "w") as f:
        f.write(proc.stdout + "\n" + proc.stderr)
    exe = os.path.join(out, "cuda_conv")
    ok = proc.returncode == 0 and os.path.exists(exe)
    print("built" if ok else "FAILED", exe)
    return 0 if ok else 2


if __name__ == "__main__":
    sys.exit(main())
