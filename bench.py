#!/usr/bin/env python
"""Headline benchmark: Mpixels/s for the 3x3 RGB Gaussian, 1920x2520, 100 iterations.

Contract (see the task statement): ``python bench.py --gpus N --steps K --warmup W`` — for N > 1
launched as ``python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...`` (one
rank per GPU, NCCL control plane). One *step* is one full job of the benchmark config: a fresh
1920x2520 RGB image filtered 100 times (BASELINE.json config 2; the image is block-decomposed over
the N GPUs, so the scaling is **strong**). W untimed warm-up steps, then exactly K timed steps
bracketed by barrier + synchronize; CUDA-event time, max over ranks; rank 0 prints ONE JSON line.

Two numbers are measured with the same engine API:

* ``value``  — device-resident: each step takes its input from a pool of distinct synthetic images
  on the device whose total size exceeds the L2 (so a step never starts from a warm cache), copies
  it into the tile buffers and iterates. Everything (copy, halo protocol, every iteration) is inside
  the timed region.
* ``e2e``    — end to end through ``ConvEngine``: every step copies its input tiles from pinned host
  memory to the device, iterates, and reads the result tiles back to pinned host memory.

``--impl reference`` runs the UNMODIFIED reference ``cuda_conv`` (built by its own Makefile into
``baseline/_ref`` by ``baseline/build_reference.py``) on the same image shape through its own CLI.
It exists only for N=1 — the reference has no multi-GPU program.
"""
from __future__ import annotations

import argparse
import json
import os
import socket
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WIDTH, HEIGHT, CHANNELS, LOOPS, FILTER = 1920, 2520, 3, 100, "gaussian"
# best published RGB rate of any reference variant (MPI+OMP, 16 processes, 20 iters,
# results/time measurements.pdf p.2 row 7). The reference's own GPU number for exactly this config
# (GTX 760, 100 iters, end to end) is 58.9 Mpx/s.
BASELINE_MPX_S = 537.6
SLOTS = 3   # buffer sets of the pipelined end-to-end path: H2D(i+1) || iterate(i) || D2H(i-1)
METRIC = "Mpixels/sec, 3x3 RGB conv 1920x2520, 100 iters (whole job, device-timed, max over ranks)"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--backend", default="fused", choices=["fused", "nccl"])
    ap.add_argument("--mode", default="auto")
    ap.add_argument("--halo-depth", type=int, default=0)
    ap.add_argument("--filter", default=FILTER)
    ap.add_argument("--no-autotune", action="store_true", help="use the model-based plan as is")
    return ap.parse_args()


def free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def relaunch_under_torchrun(a) -> int:
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={a.gpus}",
           "--master-addr", "127.0.0.1", "--master-port", str(free_port()), os.path.abspath(__file__)]
    cmd += sys.argv[1:]
    return subprocess.call(cmd)


# --------------------------------------------------------------------------------------------------
def run_ours(a) -> dict | None:
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local)
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if world != a.gpus and world > 1:
        raise SystemExit(f"--gpus {a.gpus} but WORLD_SIZE={world}")

    from parallel_convolution_b200 import ConvEngine
    from parallel_convolution_b200.ops import cuda as K
    from parallel_convolution_b200.utils.clocks import ClockSampler

    eng = ConvEngine(width=WIDTH, height=HEIGHT, channels=CHANNELS, filter=a.filter,
                     backend=a.backend, mode=a.mode, halo_depth=a.halo_depth or None, device=local,
                     slots=SLOTS if a.backend == "fused" else 1)
    tuned = {}
    if not a.no_autotune and a.backend == "fused":
        try:
            tuned = eng.autotune(LOOPS)      # untimed: pick (mode, k, R) by measurement on this box
        except Exception as e:  # noqa: BLE001 - never let the tuner take the benchmark down
            tuned = {"error": str(e)}
    tile = eng.grid.tile(rank)
    tile_bytes = tile.rows * tile.cols * CHANNELS
    l2 = K.l2_bytes() or (126 << 20)
    pool = max(4, min(160, -(-int(1.5 * l2) // tile_bytes)))   # pool bytes per GPU > L2
    eng.make_input_pool(pool, seed=1)
    stage = [eng.pinned_buffers() for _ in range(SLOTS)]   # one (input, output) staging pair per slot
    for s_i, (ins, _) in enumerate(stage):
        for r, arr in ins.items():
            arr[:] = (rank * 31 + 7 * s_i + 7) & 0xFF
    K_steps, W = a.steps, max(a.warmup, 3)

    def barrier():
        eng.synchronize()
        if dist:
            dist.barrier()

    def timed(step_fn, steps, tail_stream=None):
        """Device time from before the first enqueue to after the last operation of the last step."""
        barrier()
        n0 = K.launch_count()
        e0, e1 = K.Event(), K.Event()
        e0.record(eng.stream)
        for i in range(steps):
            step_fn(i)
        e1.record(tail_stream() if tail_stream else eng.stream)
        e1.synchronize()
        barrier()
        return e0.elapsed_ms(e1), K.launch_count() - n0

    def allmax(v):
        if not dist:
            return v
        import torch
        t = torch.tensor([v], dtype=torch.float64, device=f"cuda:{local}")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def allsum(v):
        if not dist:
            return v
        import torch
        t = torch.tensor([v], dtype=torch.float64, device=f"cuda:{local}")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    dev_step = lambda i: eng.enqueue_step_from_pool(i, LOOPS)
    if a.backend == "fused":
        e2e_step = lambda i: eng.enqueue_step_pipelined(stage[i % SLOTS][0], stage[i % SLOTS][1], LOOPS)
    else:
        e2e_step = lambda i: eng.enqueue_step_pinned(stage[0][0], stage[0][1], LOOPS)

    timed(dev_step, W)                        # warm-up (untimed)
    sampler = ClockSampler(list(range(a.gpus)) if rank == 0 else []).start() if rank == 0 else None
    ms, launches = timed(dev_step, K_steps)   # the judged number
    clocks = sampler.stop() if sampler else {}
    ms = allmax(ms)
    launches = int(allsum(launches))
    # the last operation of a pipelined step is its D2H on the copy-out stream
    tail = (lambda: eng._impl._pipeline()["s_out"]) if a.backend == "fused" else None
    timed(e2e_step, W, tail)
    ms_e2e, _ = timed(e2e_step, K_steps, tail)
    ms_e2e = allmax(ms_e2e)

    px = WIDTH * HEIGHT * LOOPS
    value = px * K_steps / (ms * 1e-3) / 1e6
    e2e = px * K_steps / (ms_e2e * 1e-3) / 1e6
    rep = eng.report
    impl = eng._impl
    kmode = impl.choose_mode(LOOPS) if hasattr(impl, "choose_mode") else "stream"
    plan = impl.plan_for(kmode, LOOPS) if hasattr(impl, "plan_for") else impl.plan(LOOPS)
    h2d = allsum(tile_bytes)
    eng.close()
    if dist:
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return None
    g = eng.grid
    from parallel_convolution_b200.utils import roofline
    kk = int(getattr(plan, "k", 1) or 1) if plan else 1
    roof = roofline.terms(WIDTH, HEIGHT, CHANNELS, LOOPS, a.gpus, kk, eng.filter.kind_name,
                          g.halo_bytes_per_iteration(0, kk))
    roof_frac = roofline.fraction(ms / K_steps * 1e-3, roof)
    return {
        "metric": METRIC, "value": round(value, 1), "unit": "Mpx/s", "n_gpus": a.gpus,
        "steps": K_steps, "warmup": W, "ms_per_step": round(ms / K_steps, 5),
        "higher_is_better": True, "scaling": "strong",
        "vs_baseline": round(value / BASELINE_MPX_S, 2),
        "dtype": "uint8 storage, integer arithmetic bit-identical to the reference's fp32 accumulate",
        "data": "synthetic random-pixel images of the benchmark shape (device generator)",
        "impl": "ours",
        "config": {"model": "3x3 gaussian blur /16 (reference default filter)", "image": f"{WIDTH}x{HEIGHT} RGB",
                   "iterations_per_step": LOOPS, "global_batch": 1, "seq_len": WIDTH * HEIGHT,
                   "parallelism": f"2-D block decomposition {g.grid_rows}x{g.grid_cols} over {a.gpus} GPU(s)",
                   "backend": a.backend, "kernel_mode": kmode,
                   "autotune": {"tried": tuned.get("tried", [])[:8]} if tuned else "off", "plan": dict(plan.__dict__) if plan else {},
                   "cache": f"inputs rotate through a device pool of {pool} images/GPU "
                            f"({pool * tile_bytes >> 20} MiB > L2 {l2 >> 20} MiB); no L2 flush needed",
                   "timed_region": "pool->tile device copy + halo protocol + all iterations of every step",
                   "roofline_s_per_step": roof, "fraction_of_binding_roofline": round(roof_frac, 4),
                   "baseline_mpx_s": BASELINE_MPX_S,
                   "baseline_source": "BASELINE.md: MPI+OMP rgb 1920x2520, 16 processes (best published RGB rate)"},
        "clocks": clocks,
        "e2e": {"value": round(e2e, 1), "unit": "Mpx/s", "ms_per_step": round(ms_e2e / K_steps, 5),
                "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(h2d),
                "path": "ConvEngine.enqueue_step_pipelined: every step copies its input tiles pinned host -> device, "
                        "iterates 100x, copies the result device -> pinned host; 3 buffer slots / 3 streams overlap "
                        "the copies of neighbouring steps with the iterations (all K steps inside the timed region)"},
        "gpu_launches": launches,
    }


# --------------------------------------------------------------------------------------------------
def run_reference(a) -> dict:
    if a.gpus != 1:
        return {"impl": "reference", "unavailable":
                "the reference has no multi-GPU program (its CUDA variant is single-device; its MPI "
                "variants are CPU-only and no MPI toolchain exists in this image)"}
    exe = os.path.join(ROOT, "baseline", "_ref", "cuda", "cuda_conv")
    if not os.path.exists(exe):
        try:
            subprocess.run([sys.executable, os.path.join(ROOT, "baseline", "build_reference.py")],
                           check=True, capture_output=True, timeout=600)
        except Exception as e:  # noqa: BLE001
            return {"impl": "reference", "unavailable": f"reference cuda_conv could not be built: {e}"}
    if not os.path.exists(exe):
        return {"impl": "reference", "unavailable": "baseline/_ref/cuda/cuda_conv missing after build"}
    import numpy as np
    work = "/tmp/pconv_ref_bench"
    os.makedirs(work, exist_ok=True)
    img = os.path.join(work, "synthetic_1920_2520.raw")
    rng = np.random.default_rng(1)
    rng.integers(0, 256, WIDTH * HEIGHT * CHANNELS, dtype=np.uint8).tofile(img)

    def once(loops: int) -> float:
        t0 = time.perf_counter()
        out = subprocess.run([exe, os.path.basename(img), str(WIDTH), str(HEIGHT), str(loops), "rgb"],
                             cwd=work, capture_output=True, text=True, timeout=900)
        wall = time.perf_counter() - t0
        for line in out.stdout.splitlines():
            if "Execution time" in line:
                return float(line.split(":")[1].split()[0])   # the reference's own timer
        raise RuntimeError(f"reference printed no time (rc={out.returncode}, wall={wall:.1f}s): {out.stderr[-300:]}")

    try:
        for _ in range(max(1, min(a.warmup, 2))):
            once(LOOPS)                        # JIT of the compute_52 PTX + page-in
        reps = max(2, min(a.steps, 4))
        long_loops = LOOPS + 1000                 # a long run isolates the per-iteration cost from the
        t100 = min(once(LOOPS) for _ in range(reps))          # ~1.3 s of start-up, file I/O and JIT
        tlong = min(once(long_loops) for _ in range(reps))
    except Exception as e:  # noqa: BLE001
        return {"impl": "reference", "unavailable": f"reference cuda_conv failed to run: {e}"}
    px = WIDTH * HEIGHT
    marginal_s = max((tlong - t100) / (long_loops - LOOPS), 1e-9)   # per-iteration device cost only
    value = px / marginal_s / 1e6
    whole = px * LOOPS / t100 / 1e6
    return {"metric": METRIC, "value": round(value, 1), "unit": "Mpx/s", "n_gpus": 1, "steps": reps,
            "warmup": max(1, min(a.warmup, 2)), "ms_per_step": round(marginal_s * LOOPS * 1e3, 4),
            "higher_is_better": True, "scaling": "strong", "vs_baseline": round(value / BASELINE_MPX_S, 3),
            "dtype": "uint8 storage, FP64 per-tap arithmetic (cuda/cuda_convolute.cu:38-40)",
            "data": "synthetic random-pixel image of the benchmark shape (raw file in /tmp)",
            "impl": "reference",
            "config": {"model": "3x3 gaussian blur /16", "image": f"{WIDTH}x{HEIGHT} RGB",
                       "iterations_per_step": LOOPS, "binary": "baseline/_ref/cuda/cuda_conv (unmodified sources, "
                       "its own Makefile: sm_52 SASS + compute_52 PTX, JIT on B200)",
                       "timing": "value = marginal loop rate from the reference's own host timer "
                                 "((t(1100 it) - t(100 it)) / 1000, best of N runs each); e2e = its "
                                 "whole-program time for 100 iterations incl. file I/O"},
            "e2e": {"value": round(whole, 1), "unit": "Mpx/s", "seconds_100_iters": t100,
                    "seconds_1100_iters": tlong},
            "gpu_launches": LOOPS}


def main() -> int:
    a = parse_args()
    if a.impl == "reference":
        if int(os.environ.get("RANK", "0")) == 0:
            print(json.dumps(run_reference(a)), flush=True)
        return 0
    if a.gpus > 1 and "WORLD_SIZE" not in os.environ:
        return relaunch_under_torchrun(a)
    out = run_ours(a)
    if out is not None:
        print(json.dumps(out), flush=True)
    return 0


if __name__ == "__main__":
    sys.exit(main())
