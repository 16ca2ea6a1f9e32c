"""Quick device-timed throughput probe (development aid; bench.py is the judged harness)."""
import argparse
import json
import sys
import os

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from parallel_convolution_b200.models.filters import Filter
from parallel_convolution_b200.ops import cuda as K
from parallel_convolution_b200.parallel.decomposition import TileGrid
from parallel_convolution_b200.parallel.fused import VirtualCluster


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--width", type=int, default=1920)
    ap.add_argument("--height", type=int, default=2520)
    ap.add_argument("--channels", type=int, default=3)
    ap.add_argument("--loops", type=int, default=100)
    ap.add_argument("--filter", default="gaussian")
    ap.add_argument("--workers", type=int, default=1)
    ap.add_argument("--k", type=int, default=0)
    ap.add_argument("--band", type=int, default=0)
    ap.add_argument("--mode", default="stream")
    ap.add_argument("--warps", type=int, default=0)
    ap.add_argument("--grid", default="")
    ap.add_argument("--reps", type=int, default=5)
    a = ap.parse_args()
    f = Filter.parse(a.filter)
    grid = tuple(int(v) for v in a.grid.split("x")) if a.grid else None
    g = TileGrid.create(a.height, a.width, a.channels, a.workers, grid)
    vc = VirtualCluster(g, f, halo_depth=a.k or None, band=a.band or None, mode=a.mode, warps=a.warps or None)
    vc.synth(1)
    plan = vc.rt.plan_for(vc.rt.choose_mode(a.loops), a.loops)
    vc.run(a.loops)
    K.device_sync()
    best = 1e9
    for _ in range(a.reps):
        e0, e1 = K.Event(), K.Event()
        e0.record(vc.rt.stream)
        vc.run(a.loops)
        e1.record(vc.rt.stream)
        e1.synchronize()
        best = min(best, e0.elapsed_ms(e1))
    st = vc.collect_status()
    mpx = a.width * a.height * a.loops / (best * 1e-3) / 1e6
    print(json.dumps({"shape": [a.height, a.width, a.channels], "filter": f.name, "path": f.kind_name,
                      "loops": a.loops, "workers": a.workers, "mode": vc.rt.choose_mode(a.loops), "plan": plan.__dict__,
                      "ms": round(best, 4), "mpx_s": round(mpx, 1),
                      "us_per_iter": round(best * 1e3 / a.loops, 3),
                      "wait_max_us": round(st.max_wait_us, 1), "wait_avg_us": round(st.sum_wait_us / max(st.waits, 1), 2),
                      "waits": st.waits, "probe": getattr(list(vc.rt.drivers.values())[0], "probe", None)}))
    vc.close()


if __name__ == "__main__":
    main()
