#!/usr/bin/env python
"""Halo-exchange sweep (BASELINE.json config 5): square tiles of edge 256..16384 px on N GPUs, one
exchange per iteration (k=1) or the planner's depth (--halo-depth 0): the fused backend (persistent kernel
with pushed halos where the tile fits on chip, streaming kernel with pulled halos otherwise) vs the NCCL
send/recv baseline (one pack kernel + one ncclGroup + one unpack kernel per exchange, whole run replayed
as a CUDA graph) vs the MPI-like host backend (--with-cpu, small edges only: `cpu` = gloo messages, `cpu:shm` = CPU
ranks over the shared image; the CPU-only sweep incl. the reference's own mpi_conv is bench/halo_sweep_cpu.py).

For every tile edge E the image is (grid_rows*E) x (grid_cols*E) grey pixels, so every rank owns
an E x E tile. Reported per backend: device time per iteration (max over ranks), the same tile's
time on ONE rank without any neighbour (measured in the same process), the difference = exposed
halo cost per exchange, and the halo bytes a rank pulls per exchange. The reference's MPI
Isend/Irecv path cannot be built here (no MPI toolchain); its stand-in is ``--backend cpu`` (the
host backend over gloo), which is included for the smaller tiles with ``--with-cpu``.

Run:  torchrun --nproc-per-node N bench/halo_sweep.py [--edges 256,1024,4096] [--loops 40]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--edges", default="256,512,1024,2048,4096,8192,16384")
    ap.add_argument("--loops", type=int, default=40)
    ap.add_argument("--halo-depth", type=int, default=1)
    ap.add_argument("--backends", default="fused:resident,fused:stream,nccl",
                    help="comma list of backend[:mode]; cpu = the MPI-like host backend over gloo")
    ap.add_argument("--with-cpu", action="store_true", help="add the gloo host backend for edges <= 1024")
    ap.add_argument("--channels", type=int, default=1)
    a = ap.parse_args()
    os.environ.setdefault("PCONV_NCCL_GRAPH", "1")     # the NCCL arm at its best: graph replay
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("cpu:gloo,cuda:nccl", device_id=torch.device("cuda", local))
    from parallel_convolution_b200 import ConvEngine
    from parallel_convolution_b200.ops import host as H
    gr, gc, _ = H.choose_grid(1 << 20, 1 << 20, world)        # square-image grid for this world size
    rank = dist.get_rank() if world > 1 else 0
    for E in [int(v) for v in a.edges.split(",")]:
        W, Hh = gc * E, gr * E
        row = {"tile_edge": E, "gpus": world, "grid": f"{gr}x{gc}", "k": a.halo_depth, "loops": a.loops}
        arms = a.backends.split(",") + (["cpu", "cpu:shm"] if a.with_cpu and E <= 1024 else [])
        done = []
        for arm in arms:
            be, _, mode = arm.partition(":")
            key = arm.replace(":", "_")
            try:
                transport = ("shm" if mode == "shm" else "gloo") if be == "cpu" else "auto"   # cpu = messages, cpu:shm = shared image
                eng = ConvEngine(width=W, height=Hh, channels=a.channels, filter="gaussian", backend=be,
                                 grid=(gr, gc), halo_depth=a.halo_depth or None,
                                 mode="auto" if be == "cpu" else (mode or "auto"), cpu_transport=transport, device=local)
                best = None
                for _ in range(3):
                    eng.load_synthetic(1)
                    rep = eng.iterate(a.loops)
                    best = rep.elapsed_ms if best is None else min(best, rep.elapsed_ms)
            except (ValueError, RuntimeError) as e:      # e.g. resident mode: the tile does not fit on chip
                row[key + "_skipped"] = str(e)[:80]
                continue
            row[key + "_us_per_iter"] = round(best * 1e3 / a.loops, 3)
            row[key + "_k"] = rep.plan.get("k", a.halo_depth)
            row["halo_bytes_per_exchange_k1"] = eng.grid.halo_bytes_per_iteration(0, 1)
            row.setdefault("waits", {})[key] = round(rep.max_wait_us, 1)
            eng.close()
            done.append(key)
        # the same tile alone on this GPU (no neighbours, no process group): the compute-only floor
        from parallel_convolution_b200.parallel.decomposition import TileGrid
        from parallel_convolution_b200.parallel.fused import VirtualCluster
        from parallel_convolution_b200.models.filters import Filter
        from parallel_convolution_b200.ops import cuda as K
        vc = VirtualCluster(TileGrid.create(E, E, a.channels, 1), Filter.builtin("gaussian"),
                            halo_depth=a.halo_depth or None, mode="auto")
        vc.synth(1)
        vc.run(a.loops)
        K.device_sync()
        solo = None
        for _ in range(3):
            e0, e1 = K.Event(), K.Event()
            e0.record(vc.rt.stream)
            vc.run(a.loops)
            e1.record(vc.rt.stream)
            e1.synchronize()
            ms = e0.elapsed_ms(e1)
            solo = ms if solo is None else min(solo, ms)
        vc.close()
        row["solo_us_per_iter"] = round(solo * 1e3 / a.loops, 3)
        for key in done:       # what the neighbours cost per iteration, on top of the same tile alone
            row[key + "_exposed_halo_us_per_iter"] = round(row[key + "_us_per_iter"] - row["solo_us_per_iter"], 3)
        if rank == 0:
            print(json.dumps(row), flush=True)
        if world > 1:
            dist.barrier()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
