#!/usr/bin/env python
"""Halo-exchange sweep on CPU ranks (BASELINE.json config 5 without a GPU): square grey tiles of edge E on N
ranks of one node, one exchange per iteration (k = 1) unless stated. Arms:

  shm_k1 / shm_k6   ranks share the image in /dev/shm, one epoch counter per rank (csrc/host/shm_ranks.cpp)
  tcp_k1 / tcp_k6   private tiles, eight packed strips per exchange over the native TCP mesh
                    (csrc/host/net_ranks.cpp): the reference's schedule (post -> interior -> wait -> border ring,
                    mpi/mpi_conv.c:164-248)
  gloo_k1           the same schedule with gloo messages driven from Python (parallel/host_backend.py)
  reference         unmodified mpi/mpi_conv.c (-O3) under baseline/mpi_shim: 4 messages per iteration

and for each the same tile alone on one rank (no neighbour): the difference is the exposed cost of an exchange.
Run under torchrun (gloo):  torchrun --nproc-per-node N bench/halo_sweep_cpu.py [--edges 256,512,1024]
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
MPI = os.path.join(ROOT, "baseline", "_ref", "mpi")


def reference_loop_s(work, w, h, loops, np_):
    import numpy as np
    name = f"sweep_{w}_{h}.raw"
    path = os.path.join(work, name)
    if not os.path.exists(path):
        np.random.default_rng(1).integers(0, 256, w * h, dtype=np.uint8).tofile(path)
    best = None
    for _ in range(3):
        p = subprocess.run([os.path.join(MPI, "mpirun_shim"), "-np", str(np_), os.path.join(MPI, "mpi_conv_O3"), name,
                            str(w), str(h), str(loops), "grey"], cwd=work, capture_output=True, text=True, timeout=600)
        if p.returncode != 0:
            return None
        t = float(p.stdout.strip().splitlines()[-1])
        best = t if best is None else min(best, t)
    return best


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--edges", default="256,512,1024,2048")
    ap.add_argument("--loops", type=int, default=60)
    a = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    import torch.distributed as dist
    if world > 1:
        dist.init_process_group("gloo")
    rank = dist.get_rank() if world > 1 else 0
    from parallel_convolution_b200 import ConvEngine
    from parallel_convolution_b200.models.filters import Filter
    from parallel_convolution_b200.ops import host as H
    gr, gc, _ = H.choose_grid(1 << 20, 1 << 20, world)
    f = Filter.builtin("gaussian")
    have_ref = os.path.exists(os.path.join(MPI, "mpi_conv_O3"))
    with tempfile.TemporaryDirectory(prefix="pconv_sweep_") as work:
        for E in [int(v) for v in a.edges.split(",")]:
            W, Hh = gc * E, gr * E
            row = {"tile_edge": E, "ranks": world, "grid": f"{gr}x{gc}", "loops": a.loops,
                   "halo_bytes_per_exchange_k1": 4 * E + 4}
            arms = (("shm_k1", "shm", 1), ("shm_k6", "shm", 6), ("tcp_k1", "tcp", 1), ("tcp_k6", "tcp", 6),
                    ("gloo_k1", "gloo", 1))
            for key, transport, k in arms:
                if world == 1:
                    break
                eng = ConvEngine(width=W, height=Hh, channels=1, filter=f, backend="cpu", grid=(gr, gc), halo_depth=k,
                                 threads=1, cpu_transport=transport, timeout_ms=120_000.0)
                best = None
                for _ in range(3):
                    eng.load_synthetic(1)
                    rep = eng.iterate(a.loops)
                    best = rep.elapsed_ms if best is None else min(best, rep.elapsed_ms)
                eng.close()
                row[key + "_us_per_iter"] = round(best * 1e3 / a.loops, 2)
            if rank == 0:
                # the same tile alone: one rank, no neighbour, the plain engine
                solo_eng = ConvEngine(width=E, height=E, channels=1, filter=f, backend="cpu", threads=1, standalone=True)
                solo = None
                for _ in range(3):
                    solo_eng.load_synthetic(1)
                    t0 = time.perf_counter()
                    solo_eng._impl.iterate(a.loops)
                    dt = time.perf_counter() - t0
                    solo = dt if solo is None else min(solo, dt)
                row["solo_us_per_iter"] = round(solo * 1e6 / a.loops, 2)
                for key, _, _ in arms:
                    if key + "_us_per_iter" in row:
                        row[key + "_exposed_us_per_iter"] = round(row[key + "_us_per_iter"] - row["solo_us_per_iter"], 2)
                if have_ref:
                    t = reference_loop_s(work, W, Hh, a.loops, world)
                    t1 = reference_loop_s(work, E, E, a.loops, 1)
                    if t and t1:
                        row["reference_us_per_iter"] = round(t * 1e6 / a.loops, 2)
                        row["reference_solo_us_per_iter"] = round(t1 * 1e6 / a.loops, 2)
                        row["reference_exposed_us_per_iter"] = round((t - t1) * 1e6 / a.loops, 2)
                print(json.dumps(row), flush=True)
            if world > 1:
                dist.barrier()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
