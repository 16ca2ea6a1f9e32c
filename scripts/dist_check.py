"""Multi-process correctness check (run under torchrun): every backend vs the CPU engine.

    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 scripts/dist_check.py
      [--backend fused|nccl|cpu] [--height H --width W --channels C --loops L --filter F]
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--backend", default="fused")
    ap.add_argument("--height", type=int, default=480)
    ap.add_argument("--width", type=int, default=640)
    ap.add_argument("--channels", type=int, default=3)
    ap.add_argument("--loops", type=int, default=12)
    ap.add_argument("--filter", default="gaussian")
    ap.add_argument("--halo-depth", type=int, default=0)
    ap.add_argument("--reps", type=int, default=2)
    ap.add_argument("--mode", default="auto")
    ap.add_argument("--grid", default="")
    ap.add_argument("--warps", type=int, default=0)
    a = ap.parse_args()
    import torch
    import torch.distributed as dist
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if a.backend == "cpu":
        dist.init_process_group("gloo")
    else:
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from parallel_convolution_b200 import ConvEngine, Filter
    from parallel_convolution_b200.ops import host
    rank = dist.get_rank()
    f = Filter.parse(a.filter)
    g_rows = host.synth_rect(42, a.width * a.channels, 0, a.height, 0, a.width * a.channels)
    img = g_rows.reshape(a.height, a.width, a.channels) if a.channels > 1 else g_rows
    ref = host.cpu_run_image(img, f.spec, a.loops) if rank == 0 else None
    grid = tuple(int(v) for v in a.grid.split("x")) if a.grid else None
    eng = ConvEngine(width=a.width, height=a.height, channels=a.channels, filter=f, backend=a.backend,
                     halo_depth=a.halo_depth or None, mode=a.mode, grid=grid,
                     **({"warps": a.warps} if a.warps else {}))
    ok = True
    for rep in range(a.reps):     # twice: exercises epoch continuity / buffer reuse across runs
        eng.load_synthetic(42)
        r = eng.iterate(a.loops)
        out = eng.gather()
        if rank == 0:
            same = np.array_equal(out.reshape(ref.shape), ref)
            ok &= same
            print(f"[dist_check] rep={rep} backend={a.backend} world={dist.get_world_size()} "
                  f"grid={eng.grid.grid_rows}x{eng.grid.grid_cols} filter={f.name} loops={a.loops} "
                  f"elapsed_ms={r.elapsed_ms:.3f} max_wait_us={r.max_wait_us:.1f} waits={r.waits} "
                  f"mode={eng._impl.choose_mode(a.loops) if hasattr(eng._impl, 'choose_mode') else '-'} "
                  f"plan={r.plan} probe={r.probe} {'OK' if same else 'MISMATCH ' + str(int((out.reshape(ref.shape) != ref).sum()))}",
                  flush=True)
    eng.close()
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0:
        print("DIST_CHECK", "OK" if ok else "FAILED", flush=True)
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
