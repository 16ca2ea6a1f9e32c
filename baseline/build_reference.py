#!/usr/bin/env python
"""Build the UNMODIFIED reference into baseline/_ref/ (git-ignored; travels with gpurun).

The reference is not a Python package (no setup.py / pyproject.toml), so the prescribed
``pip install --no-index ... --target baseline/_ref /root/reference`` cannot work; the outcome of
that attempt is recorded in DESIGN.md. What *can* be built offline is its CUDA program: its sources
are copied verbatim from /root/reference/cuda and compiled with its own Makefile (no -arch: nvcc
12.9 embeds sm_52 SASS + compute_52 PTX, which the driver JIT-compiles on a B200). The MPI variants
need mpicc/mpi.h, which this image does not have: they are compiled, unmodified, against the
single-node MPI stand-in of baseline/mpi_shim (see ``build_mpi``).
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
    # Second bar: the same unmodified sources with the one flag the shipped Makefile lacks,
    # -arch=sm_100a on its nvcc compile line (cuda/Makefile:11) — native Blackwell SASS instead of a JIT of
    # the compute_52 PTX. The Makefile's three compile commands and its link command, replayed verbatim.
    out2 = os.path.join(DST, "cuda_sm100a")
    if os.path.isdir(out2):
        writable(out2)
        shutil.rmtree(out2)
    shutil.copytree(src, out2)
    writable(out2)
    log2 = []
    ok2 = True
    for cmd in (["gcc", "-c", "main.c", "-O2"], ["gcc", "-c", "funcs.c", "-O2"],
                ["nvcc", "-c", "cuda_convolute.cu", "-arch=sm_100a"],
                ["nvcc", "-o", "cuda_conv", "main.o", "funcs.o", "cuda_convolute.o", "-O2"]):
        p2 = subprocess.run(cmd, cwd=out2, env=env, capture_output=True, text=True)
        log2.append(" ".join(cmd) + "\n" + p2.stdout + p2.stderr)
        ok2 = ok2 and p2.returncode == 0
    with open(os.path.join(DST, "build_sm100a.log"), "w") as f:
        f.write("\n".join(log2))
    print("built" if ok2 else "FAILED", os.path.join(out2, "cuda_conv"))
    build_mpi()
    return 0 if ok else 2


def build_mpi() -> bool:
    """The CPU programs: unmodified mpi/mpi_conv.c and mpi_omp/mpi_omp_conv.c, compiled against the single-node
    MPI stand-in of baseline/mpi_shim (this image has no mpicc / mpi.h / libmpi). Two builds of mpi_conv: the
    documented command line (`mpicc -o mpi_conv mpi_conv.c`, mpi/READ_ME.txt:2 — no optimisation flag) and the
    same with -O3. mpi_omp_conv does not compile with -fopenmp (collapse(3) on a 2-deep nest; the error goes to
    the log) and the documented `-openmp` means `-o penmp` to gcc, so it is built without OpenMP, as a GNU user
    of the reference would end up with. The binaries are portable x86-64 (no -march), they travel to the GPU box.
    """
    shim = os.path.join(ROOT, "baseline", "mpi_shim")
    out = os.path.join(DST, "mpi")
    if os.path.isdir(out):
        shutil.rmtree(out)
    os.makedirs(out)
    log, ok = [], True

    def run(cmd, must=True):
        nonlocal ok
        p = subprocess.run(cmd, cwd=out, capture_output=True, text=True)
        log.append(" ".join(cmd) + f"\n[rc={p.returncode}]\n" + p.stdout + p.stderr)
        if must:
            ok = ok and p.returncode == 0
        return p.returncode == 0

    for rel in ("mpi/mpi_conv.c", "mpi_omp/mpi_omp_conv.c"):
        src = os.path.join(REF, rel)
        if not os.path.exists(src):
            print(f"reference source not found: {src}", file=sys.stderr)
            return False
        dst = os.path.join(out, os.path.basename(rel))
        shutil.copyfile(src, dst)              # verbatim; only the MPI library underneath is a stand-in
        os.chmod(dst, 0o644)
    inc, lib = "-I" + shim, os.path.join(shim, "mpi_shim.c")
    run(["gcc", "-O2", "-o", "mpirun_shim", os.path.join(shim, "mpirun_shim.c")])
    run(["gcc", "-O2", inc, "-o", "selftest", os.path.join(shim, "selftest.c"), lib])
    run(["gcc", inc, "-o", "mpi_conv", "mpi_conv.c", lib])                    # as documented
    run(["gcc", "-O3", inc, "-o", "mpi_conv_O3", "mpi_conv.c", lib])
    run(["gcc", "-O3", "-fopenmp", inc, "-o", "mpi_omp_conv_fopenmp", "mpi_omp_conv.c", lib], must=False)
    run(["gcc", "-O3", inc, "-o", "mpi_omp_conv", "mpi_omp_conv.c", lib])
    with open(os.path.join(DST, "build_mpi.log"), "w") as f:
        f.write("\n".join(log))
    print("built" if ok else "FAILED", os.path.join(out, "mpi_conv"), "(+ mpi_conv_O3, mpi_omp_conv, mpirun_shim)")
    return ok


if __name__ == "__main__":
    sys.exit(main())
