"""Per-CTA phase times of one persistent-kernel run (development aid): which CTAs wait least, i.e. pace
the grid. Every CTA leaves (push, signal+wait+fetch, compute) ns in words 10..15 of its flag line."""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from parallel_convolution_b200.models.filters import Filter
from parallel_convolution_b200.ops import cuda as K
from parallel_convolution_b200.parallel.decomposition import TileGrid
from parallel_convolution_b200.parallel.fused import VirtualCluster


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--warps", type=int, default=16)
    ap.add_argument("--k", type=int, default=5)
    ap.add_argument("--band", type=int, default=14)
    ap.add_argument("--loops", type=int, default=100)
    a = ap.parse_args()
    g = TileGrid.create(2520, 1920, 3, 1)
    vc = VirtualCluster(g, Filter.parse("gaussian"), halo_depth=a.k, band=a.band, mode="resident", warps=a.warps)
    vc.synth(1)
    for _ in range(3):
        vc.run(a.loops)
    K.device_sync()
    plan = vc.rt.resident_plan(a.loops)
    n = plan.nty * plan.ntx
    raw = np.zeros(n * 16, dtype=np.uint32)
    K.memcpy_async(raw.ctypes.data, vc.rt.mems[0].arena, n * 64, K.D2H, vc.rt.stream)
    K.stream_sync(vc.rt.stream)
    w = raw.reshape(n, 16)[:, 10:16].copy().view(np.uint64).reshape(n, 3).astype(np.float64) / 1e3
    for name, col in (("push_us", 0), ("wait_us", 1), ("compute_us", 2)):
        print(name)
        grid = w[:, col].reshape(plan.nty, plan.ntx)
        for row in grid:
            print("  " + " ".join(f"{v:6.1f}" for v in row))
    tot = w.sum(axis=1).reshape(plan.nty, plan.ntx)
    print("sum_us min/median/max", tot.min(), np.median(tot), tot.max())
    vc.close()


if __name__ == "__main__":
    main()
