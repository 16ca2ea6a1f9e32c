#!/usr/bin/env python
"""Filters are a runtime choice (the reference selects its filter by editing the source, mpi/mpi_conv.c:103-105).

Three ways to name one, and which exact arithmetic path each takes:
  * a built-in            Filter.builtin("sharpen")
  * integers + divisor    Filter.from_int([1,2,1, 2,4,2, 1,2,1], 16)     the reference's own construction
  * arbitrary fp32 taps   Filter.custom([...])
Every path reproduces the reference's single-process semantics bit for bit: fp32 accumulation in tap order,
truncation, saturation to [0, 255] (the oracle, csrc/host/oracle.cpp).
"""
import numpy as np

from parallel_convolution_b200 import ConvEngine, Filter, available_filters
from parallel_convolution_b200.ops import host


def main():
    rng = np.random.default_rng(0)
    img = rng.integers(0, 256, (96, 128, 3), dtype=np.uint8)
    print("built-in filters:", ", ".join(available_filters()))
    filters = [Filter.builtin("sharpen"),
               Filter.from_int([1, 2, 1, 2, 4, 2, 1, 2, 1], 16, name="gaussian-by-hand"),
               Filter.from_int([1, 1, 1, 1, 1, 1, 1, 1, 1], 9, name="box"),
               Filter.custom([0.05, 0.1, 0.05, 0.1, 0.4, 0.1, 0.05, 0.1, 0.05], name="soft")]
    for f in filters:
        with ConvEngine(width=128, height=96, channels=3, filter=f) as eng:
            out = eng.run(img, loops=5)
        exact = np.array_equal(out, host.oracle(img, f.spec, 5).reshape(out.shape))
        print(f"{f.name:18s} path={f.kind_name:12s} mean {img.mean():6.2f} -> {out.mean():6.2f}   equals the oracle: {exact}")
        assert exact


if __name__ == "__main__":
    main()
