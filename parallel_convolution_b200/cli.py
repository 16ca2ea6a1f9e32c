"""Command line: drop-in for the reference programs, plus the options they never had.

Reference grammar (``mpi/mpi_conv.c:340-360``, ``cuda/funcs.c:10-29``), kept verbatim::

    pconv image width height loops {rgb|grey}

* single process:           ``python -m parallel_convolution_b200 img.raw 1920 2520 50 grey``
  (the reference's ``./cuda_conv ...``; add ``--backend cpu`` for a GPU-less box)
* one process per GPU:       ``torchrun --standalone --local-addr 127.0.0.1 --nproc-per-node 8 -m
  parallel_convolution_b200 img.raw 1920 2520 50 rgb`` (the reference's ``mpirun -np 8 ./mpi_conv``)
* CPU ranks:                 same with ``--backend cpu`` (``mpi_conv`` / ``mpi_omp_conv --threads T``); ranks of one
                             node share the image in /dev/shm and trade no messages (``--cpu-transport shm``),
                             ranks of several nodes exchange gloo messages

Output: ``blur_<basename>`` next to the input (``-o`` overrides) and the elapsed seconds of the
iteration loop on stdout, max over ranks (``mpi/mpi_conv.c:273-283`` prints ``%f\\n``); ``--json``
prints a machine-readable report instead. Errors are reported with a message and a non-zero exit,
never by hanging (the reference spins forever on a short file, ``cuda/funcs.c:39-45``).
"""
from __future__ import annotations

import argparse
import json
import os
import sys

from .utils import checkpoint as ckpt
from .utils.nvtx import nvtx_range

USAGE_HINT = "image_name width height loops [rgb/grey]"


def build_parser() -> argparse.ArgumentParser:
    p = argparse.ArgumentParser(prog="pconv", description=__doc__,
                                formatter_class=argparse.RawDescriptionHelpFormatter)
    p.add_argument("image", help="raw headerless image (or any name with --synthetic)")
    p.add_argument("width", type=int)
    p.add_argument("height", type=int)
    p.add_argument("loops", type=int)
    p.add_argument("color", choices=["rgb", "grey", "gray", "rgba"],
                   help="rgb / grey as in the reference; rgba (4 interleaved channels) is an extension")
    p.add_argument("--filter", default="gaussian",
                   help="builtin name (box, gaussian, edge, sharpen, laplacian, emboss, sobel_x, sobel_y, "
                        "identity), 'a,b,c,d,e,f,g,h,i[/div]', nine floats, or a file")
    p.add_argument("--backend", default="auto", choices=["auto", "fused", "nccl", "cpu", "host"])
    p.add_argument("--mode", default="auto", choices=["auto", "stream", "resident"],
                   help="fused backend kernel: resident = ONE persistent launch per run, the tile stays in "
                        "registers and CTAs push their halos into each other's inboxes (needs the tile to "
                        "fit on chip); stream = k fused iterations per launch, any image size; auto = "
                        "resident when it fits")
    p.add_argument("--warps", type=int, default=0, choices=[0, 8, 16],
                   help="persistent kernel: warps per CTA (8 = two CTAs per SM hide each other's exchanges, "
                        "16 = one CTA per SM; 0 = planner / autotuner decides)")
    p.add_argument("--gpus", "--workers", dest="workers", type=int, default=None,
                   help="single process: decompose over N virtual ranks on one GPU")
    p.add_argument("-np", "--np", dest="nprocs", type=int, default=None, metavar="N",
                   help="spawn N ranks on this node (one per GPU, or CPU ranks with --backend cpu): the "
                        "reference's `mpirun -np N`; equivalent to launching this module under torchrun")
    p.add_argument("--grid", default=None, help="tile grid ROWSxCOLS (default: perimeter-minimising)")
    p.add_argument("--halo-depth", type=int, default=None, help="temporal-blocking depth k")
    p.add_argument("--autotune", action="store_true",
                   help="fused backend: measure candidate launch plans on this machine before the run")
    p.add_argument("--threads", type=int, default=0, help="cpu backend: OpenMP threads per rank")
    p.add_argument("--cpu-transport", default="auto", choices=["auto", "shm", "tcp", "gloo"],
                   help="cpu backend with several ranks: halos through a shared image in /dev/shm (one node; auto "
                        "picks it when possible), as messages over the native TCP mesh (any nodes; auto otherwise) "
                        "or as gloo messages")
    p.add_argument("--out-of-core", type=int, default=None, metavar="MiB",
                   help="cpu backend, single process: stream the file through row strips with k-deep halos, "
                        "holding at most ~MiB of the image in memory (images larger than RAM)")
    p.add_argument("--batch", default=None, metavar="LIST",
                   help="text file with one raw image path per line (all of the given shape): filter them "
                        "all, streaming through the pipelined slots; outputs blur_<name> (in -o DIR if given)")
    p.add_argument("-o", "--output", default=None)
    p.add_argument("--no-output", action="store_true", help="skip writing the result image")
    p.add_argument("--synthetic", type=int, default=None, metavar="SEED",
                   help="ignore the input file; generate a random image per tile")
    p.add_argument("--json", action="store_true", help="print a JSON report instead of the seconds")
    p.add_argument("--checkpoint-every", type=int, default=0, metavar="K",
                   help="write the state image + JSON sidecar every K iterations")
    p.add_argument("--resume", action="store_true", help="continue from the checkpoint of this output")
    p.add_argument("--verify", action="store_true",
                   help="recompute on the CPU oracle engine and compare byte for byte (small images)")
    p.add_argument("--exact-ref", action="store_true",
                   help="accepted for compatibility: every backend is ALWAYS bit-identical to the reference's "
                        "single-process fp32 semantics (there is no approximate mode to switch off)")
    p.add_argument("--timeout-ms", type=float, default=10_000.0, help="in-kernel halo watchdog")
    p.add_argument("--inject-drop-signal", type=int, default=None, metavar="RANK",
                   help="fault injection: RANK never signals its neighbours (watchdog demo)")
    return p


def _init_dist(backend: str):
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world <= 1:
        return None
    import datetime
    import torch
    import torch.distributed as dist
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    # a rank that dies (bad file, watchdog) must not leave the others in a collective for 10 minutes
    limit = datetime.timedelta(seconds=int(os.environ.get("PCONV_DIST_TIMEOUT_S", "180")))
    if backend in ("cpu", "host") or not torch.cuda.is_available():
        dist.init_process_group("gloo", timeout=limit)
    else:
        local = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local), timeout=limit)
    return dist


def _spawn_ranks(nprocs: int, argv: list[str]) -> int:
    """``pconv -np N ...``: re-launch this command as N ranks under ``torch.distributed.run``.

    The reference is started as ``mpirun -np N ./mpi_conv ...`` (``mpi/READ_ME.txt:5``); here the launcher
    is torchrun with a loopback rendezvous, so the whole thing stays one command."""
    import subprocess
    rest, skip = [], False
    for tok in argv:
        if skip:
            skip = False
        elif tok in ("-np", "--np"):
            skip = True
        elif not tok.startswith(("--np=", "-np=")):
            rest.append(tok)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--standalone", "--local-addr", "127.0.0.1",
           "--nnodes", "1", "--nproc-per-node", str(nprocs), "-m", __package__] + rest
    return subprocess.call(cmd)


def main(argv: list[str] | None = None) -> int:
    parser = build_parser()
    if argv is None:
        argv = sys.argv[1:]
    try:
        a = parser.parse_args(argv)
    except SystemExit as e:
        if e.code not in (0, None):
            sys.stderr.write(f"\nError Input!\n{parser.prog} {USAGE_HINT}\n\n")
        raise
    if a.width <= 0 or a.height <= 0 or a.loops < 0:
        sys.stderr.write("width and height must be positive, loops non-negative\n")
        return 2
    if a.nprocs is not None and int(os.environ.get("WORLD_SIZE", "1")) <= 1:
        if a.nprocs < 1:
            sys.stderr.write("-np needs a positive rank count\n")
            return 2
        if a.nprocs > 1:
            return _spawn_ranks(a.nprocs, list(argv))
    channels = {"rgb": 3, "rgba": 4}.get(a.color, 1)
    if a.out_of_core is not None:
        return _out_of_core(a, channels)
    grid = tuple(int(v) for v in a.grid.lower().split("x")) if a.grid else None
    dist = _init_dist(a.backend)
    rank = dist.get_rank() if dist else 0

    from .engine import ConvEngine
    import time as _time
    t_job = _time.perf_counter()
    t_load = t_store = 0.0
    out_path = a.output
    if out_path is None:
        d, b = os.path.split(a.image)
        out_path = os.path.join(d, "blur_" + b)
    try:
        with nvtx_range("pconv:setup"):
            eng = ConvEngine(width=a.width, height=a.height, channels=channels, filter=a.filter,
                             backend=a.backend, workers=a.workers, grid=grid, halo_depth=a.halo_depth,
                             mode=a.mode, warps=a.warps or None, threads=a.threads, cpu_transport=a.cpu_transport,
                             timeout_ms=a.timeout_ms,
                             slots=3 if a.batch else 1)
        if a.inject_drop_signal is not None and hasattr(eng._impl, "drivers"):
            for r, d in eng._impl.drivers.items():
                d.inject_drop_signal = (r == a.inject_drop_signal)
        if a.batch:
            with open(a.batch) as f:
                paths = [ln.strip() for ln in f if ln.strip()]
            info = eng.run_batch(paths, a.loops, a.output)
            if rank == 0:
                print(json.dumps(info) if a.json else f"{info['seconds']:f}")
            eng.close()
            if dist:
                dist.barrier()
                dist.destroy_process_group()
            return 0
        done = 0
        if a.autotune and a.loops > 0:
            with nvtx_range("pconv:autotune"):
                eng.autotune(min(a.loops, a.checkpoint_every) if a.checkpoint_every > 0 else a.loops)
        with nvtx_range("pconv:load"):
            t0_load = _time.perf_counter()
            state = ckpt.load_state(out_path) if a.resume else None
            if state is not None:
                ckpt.check_compatible(state, a.width, a.height, channels, eng.filter, a.loops)
                eng.load_file(ckpt.image_path(out_path, state))
                done = state["iterations_done"]
            elif a.synthetic is not None:
                eng.load_synthetic(a.synthetic)
            else:
                eng.load_file(a.image)
            t_load = _time.perf_counter() - t0_load
        elapsed_ms, launches, rep = 0.0, 0, None
        with nvtx_range("pconv:iterate"):
            while done < a.loops or rep is None:
                n = a.loops - done
                if a.checkpoint_every > 0:
                    n = min(n, a.checkpoint_every)
                rep = eng.iterate(n)
                elapsed_ms += rep.elapsed_ms
                launches += rep.launches
                done += n
                if a.checkpoint_every > 0 and done < a.loops:
                    ckpt.save_state(eng, out_path, done, a)
                if n == 0:
                    break
        if not a.no_output:
            with nvtx_range("pconv:store"):
                t0_store = _time.perf_counter()
                eng.save_file(out_path)
                t_store = _time.perf_counter() - t0_store
                if rank == 0:
                    ckpt.clear_state(out_path)
        verified = None
        if a.verify:
            verified = _verify(eng, a, channels, out_path)
        if rank == 0:
            if a.json:
                from .utils import roofline
                g = eng.grid
                mpx = a.width * a.height * a.loops / max(elapsed_ms, 1e-9) / 1e3
                kk = rep.plan.get("k", 1) or 1
                roof = roofline.terms(a.width, a.height, channels, a.loops, g.workers, kk,
                                      eng.filter.kind_name, g.halo_bytes_per_iteration(0, kk))
                print(json.dumps({
                    "seconds": elapsed_ms / 1e3, "mpixels_per_s": mpx, "iterations": a.loops,
                    "backend": eng.backend, "filter": eng.filter.describe(), "grid": g.describe(),
                    "workers": g.workers, "plan": rep.plan, "kernel_launches": launches,
                    "kernel_mode": (eng._impl.choose_mode(a.loops) if hasattr(eng._impl, "choose_mode") and a.loops else None),
                    "exchange_probe_us": rep.probe or None,
                    # whole job on this rank's clock: setup + rank-parallel read (pread overlapped with H2D) +
                    # iterations + rank-parallel write (D2H overlapped with pwrite), everything but process start
                    "wall_seconds": {"total": _time.perf_counter() - t_job, "load": t_load, "store": t_store},
                    "exposed_halo_wait_us": {"max": rep.max_wait_us, "sum": rep.sum_wait_us, "waits": rep.waits},
                    "halo_bytes_per_exchange_rank0": g.halo_bytes_per_iteration(0, rep.plan.get("k", 1) or 1),
                    "roofline": roof, "fraction_of_binding_roofline": round(
                        roofline.fraction(elapsed_ms / 1e3, roof), 4) if a.loops else None,
                    "output": None if a.no_output else out_path, "verified": verified}))
            else:
                print(f"{elapsed_ms / 1e3:f}")
        eng.close()
    except (OSError, ValueError, RuntimeError) as e:
        sys.stderr.write(f"pconv: {e}\n")
        if dist:
            try:
                dist.destroy_process_group()
            except Exception:
                pass
        return 1
    if dist:
        dist.barrier()
        dist.destroy_process_group()
    return 0 if verified in (None, True) else 3


def _out_of_core(a, channels: int) -> int:
    """``--out-of-core MiB``: file -> file on the native CPU engine without ever loading the image."""
    import time
    from .models.filters import Filter
    from .ops import host as H
    gpu = a.backend in ("fused", "gpu", "cuda")
    if a.backend == "auto":
        from .ops import cuda as K
        gpu = K.available()
    if gpu:
        return _out_of_core_gpu(a, channels)
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        sys.stderr.write("pconv: --out-of-core on the cpu backend runs in a single process\n")
        return 2
    out_path = a.output
    if out_path is None:
        d, b = os.path.split(a.image)
        out_path = os.path.join(d, "blur_" + b)
    try:
        filt = Filter.parse(a.filter)
        t0 = time.perf_counter()
        H.cpu_run_file(a.image, out_path, a.height, a.width, channels, filt.spec, a.loops, a.threads,
                       a.halo_depth or 0, max(1, a.out_of_core) << 20)
        dt = time.perf_counter() - t0
    except (OSError, ValueError, RuntimeError) as e:
        sys.stderr.write(f"pconv: {e}\n")
        return 1
    if a.json:
        print(json.dumps({"seconds": dt, "iterations": a.loops, "backend": "cpu (out of core)",
                          "mpixels_per_s": a.width * a.height * a.loops / max(dt, 1e-9) / 1e6,
                          "memory_budget_mib": a.out_of_core, "filter": filt.describe(), "output": out_path}))
    else:
        print(f"{dt:f}")
    return 0


def _out_of_core_gpu(a, channels: int) -> int:
    """``--out-of-core MiB`` on the GPU: strips of the file stream through the device with k-deep halos,
    file I/O, H2D, iterations and D2H overlapped; with ``-np N`` every rank takes every N-th strip
    (utils/out_of_core.py)."""
    from .models.filters import Filter
    from .utils import out_of_core as ooc
    dist = _init_dist("fused")
    rank = dist.get_rank() if dist else 0
    world = dist.get_world_size() if dist else 1
    out_path = a.output
    if out_path is None:
        d, b = os.path.split(a.image)
        out_path = os.path.join(d, "blur_" + b)
    try:
        filt = Filter.parse(a.filter)
        local = int(os.environ.get("LOCAL_RANK", "0"))
        info = ooc.run_file(a.image, out_path, a.width, a.height, channels, filt, a.loops,
                            max(1, a.out_of_core) << 20, k=a.halo_depth or None, device=local, rank=rank,
                            world=world, dist=dist)
    except (OSError, ValueError, RuntimeError) as e:
        sys.stderr.write(f"pconv: {e}\n")
        return 1
    finally:
        if dist:
            dist.barrier()
            dist.destroy_process_group()
    if rank == 0:
        if a.json:
            info.update({"iterations": a.loops, "backend": "fused (gpu out of core)", "memory_budget_mib": a.out_of_core,
                         "mpixels_per_s": a.width * a.height * a.loops / max(info["seconds"], 1e-9) / 1e6,
                         "filter": filt.describe(), "output": out_path})
            print(json.dumps(info))
        else:
            print(f"{info['seconds']:f}")
    return 0


def _verify(eng, a, channels: int, out_path: str) -> bool | None:
    """Rank 0 recomputes the whole image on the CPU engine and compares with the gathered result."""
    import numpy as np
    from .ops import host as H
    out = eng.gather()
    if eng.rank != 0:
        return None
    rb = a.width * channels
    if a.synthetic is not None:
        src = H.synth_rect(a.synthetic, rb, 0, a.height, 0, rb)
    else:
        src = H.read_rect(ckpt.original_input(a.image), rb, 0, a.height, 0, rb)
    img = src.reshape(a.height, a.width, channels) if channels > 1 else src
    ref = H.cpu_run_image(img, eng.filter.spec, a.loops, a.threads)
    same = bool(np.array_equal(ref.reshape(out.shape), out))
    if not same:
        sys.stderr.write(f"pconv: VERIFY FAILED: {(ref.reshape(out.shape) != out).sum()} bytes differ\n")
    return same


if __name__ == "__main__":
    sys.exit(main())
