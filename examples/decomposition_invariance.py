#!/usr/bin/env python
"""The same image on 1, 2, 4 and 6 tiles gives the same bytes.

On a GPU the tiles are "virtual ranks" of one device (the multi-GPU kernels and their halo protocol with peer
pointers that happen to be local); on the CPU backend a single process has one tile, so there the check runs over
halo depths instead. Real multi-process runs: `python -m parallel_convolution_b200 -np N ...`.
The reference's own np > 1 output differs from its np = 1 output around the tile cuts
(scripts/reference_crosscheck_mpi.py).
"""
import numpy as np

from parallel_convolution_b200 import ConvEngine
from parallel_convolution_b200.ops import cuda as K


def main():
    img = np.random.default_rng(1).integers(0, 256, (210, 331, 3), dtype=np.uint8)   # nothing divides evenly
    gpu = K.available()
    want = None
    for workers in (1, 2, 4, 6):
        for depth in (None, 3):
            kw = dict(workers=workers) if gpu else {}
            with ConvEngine(width=331, height=210, channels=3, filter="gaussian", halo_depth=depth, **kw) as eng:
                out = eng.run(img, loops=9)
                grid = f"{eng.grid.grid_rows}x{eng.grid.grid_cols}"
            want = out if want is None else want
            same = np.array_equal(out, want)
            print(f"workers={workers} grid={grid} halo depth={depth or 'auto'}: identical to the first run: {same}")
            assert same
        if not gpu:
            break


if __name__ == "__main__":
    main()
