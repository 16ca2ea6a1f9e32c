#!/usr/bin/env python
"""The reference's published CPU tables (MPI, 20 iterations, grey/rgb x 4 image heights x
process counts) re-measured on THIS box, next to this framework's CPU backend on the same cells.

  reference   unmodified mpi/mpi_conv.c, gcc -O3, under the single-node MPI stand-in of baseline/mpi_shim
              (baseline/_ref/mpi, built by baseline/build_reference.py); its own timer = iteration loop
              only, max over ranks (mpi/mpi_conv.c:162,250,273-283)
  ours ranks  `pconv -np N --backend cpu --threads 1`: N processes over the shared image (shm transport),
              or gloo messages with --transport gloo; the engine's timer = iteration loop, max over ranks
  ours omp    one process, N OpenMP threads (the mpi_omp_conv reading of "N workers")

Prints a markdown table (and writes it with --out). Images are synthetic random bytes of the named shapes.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
MPI = os.path.join(ROOT, "baseline", "_ref", "mpi")


def reference(work, name, w, h, iters, color, np_, reps):
    exe, run = os.path.join(MPI, "mpi_conv_O3"), os.path.join(MPI, "mpirun_shim")
    best = None
    for _ in range(reps):
        p = subprocess.run([run, "-np", str(np_), exe, name, str(w), str(h), str(iters), color], cwd=work,
                           capture_output=True, text=True, timeout=900)
        if p.returncode != 0:
            return None           # e.g. no exact division for this process count (mpi/mpi_conv.c:56-61)
        t = float(p.stdout.strip().splitlines()[-1])
        best = t if best is None else min(best, t)
    return best


def ours(work, name, w, h, iters, color, np_, threads, transport, reps):
    best = None
    for _ in range(reps):
        cmd = [sys.executable, "-m", "parallel_convolution_b200", name, str(w), str(h), str(iters), color,
               "--backend", "cpu", "--threads", str(threads), "--no-output", "--json", "--cpu-transport", transport]
        if np_ > 1:
            cmd[3:3] = ["-np", str(np_)]
        env = dict(os.environ, PYTHONPATH=ROOT, OMP_NUM_THREADS=str(threads))
        p = subprocess.run(cmd, cwd=work, capture_output=True, text=True, timeout=900, env=env)
        if p.returncode != 0:
            raise RuntimeError(p.stderr[-500:])
        t = json.loads(p.stdout.strip().splitlines()[-1])["seconds"]
        best = t if best is None else min(best, t)
    return best


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--np", default="1,2,4,8")
    ap.add_argument("--heights", default="630,1260,2520,5040")
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--transport", default="shm")
    ap.add_argument("--out", default="")
    a = ap.parse_args()
    import numpy as np
    nps = [int(x) for x in a.np.split(",") if int(x) <= (os.cpu_count() or 1)]
    W = 1920
    lines = [f"CPU backend vs the reference's `mpi_conv` on the same box ({os.cpu_count()} cores), {a.iters} iterations, "
             f"Gaussian /16; seconds of the iteration loop, best of {a.reps}, max over ranks.", "",
             "| image | program | " + " | ".join(f"{n} worker{'s' if n > 1 else ''}" for n in nps) + " |",
             "|---|---|" + "---|" * len(nps)]
    with tempfile.TemporaryDirectory(prefix="pconv_cpu_tables_") as work:
        for color, C in (("grey", 1), ("rgb", 3)):
            for h in [int(x) for x in a.heights.split(",")]:
                name = f"synthetic_{color}_{W}_{h}.raw"
                np.random.default_rng(h + C).integers(0, 256, W * h * C, dtype=np.uint8).tofile(os.path.join(work, name))
                rows = {"reference `mpi_conv` -O3": [reference(work, name, W, h, a.iters, color, n, a.reps) for n in nps],
                        f"ours, N ranks ({a.transport})": [ours(work, name, W, h, a.iters, color, n, 1, a.transport, a.reps) for n in nps],
                        "ours, 1 rank x N threads": [ours(work, name, W, h, a.iters, color, 1, n, a.transport, a.reps) for n in nps]}
                for prog, ts in rows.items():
                    cells = ["—" if t is None else f"{t:.4f} s ({W * h * a.iters / t / 1e6:.0f} Mpx/s)" for t in ts]
                    lines.append(f"| {color} {W}x{h} | {prog} | " + " | ".join(cells) + " |")
                    print(lines[-1], flush=True)
    text = "\n".join(lines) + "\n"
    if a.out:
        with open(a.out, "w") as f:
            f.write(text)
    else:
        print(text)


if __name__ == "__main__":
    main()
