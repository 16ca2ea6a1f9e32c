"""GPU out-of-core mode: images larger than the (aggregate) device memory, file -> file.

The image never exists in memory as a whole. It is cut into horizontal *strips* of ``S`` rows; one
pass advances every strip by ``k`` iterations: a strip is read with a ``k``-row halo above and below
(``S + 2k`` rows, clipped at the true image border), filtered ``k`` times on the GPU as an image of
its own, and its ``S`` interior rows are written to the pass output. The halo rows go stale one row
per iteration from the cut, so the interior is bit-identical to ``k`` global passes — the same
temporal-blocking argument the kernels use on chip, one level up (the CPU engine does the same in
``pconv_cpu_run_file``). ``ceil(loops / k)`` passes ping-pong between two files.

Everything overlaps: the host ``pread``s strip *i+1* into pinned memory and writer threads ``pwrite``
strip *i-1* from pinned memory while the copy engines move strip *i* (H2D / D2H on their own streams,
three buffer slots, per-slot completion events) and the SMs iterate — ``ConvEngine.enqueue_step_pipelined``.

Multi-rank (``pconv -np N ... --out-of-core MiB``): strips are independent inside a pass, so rank
``r`` simply takes strips ``r, r+N, r+2N, ...`` with a private single-GPU engine and its own file
reads/writes — rank-parallel I/O like the reference's per-rank ``MPI_File_seek`` + read/write loops
(``mpi/mpi_conv.c:133-147,253-271``), with one barrier per pass instead of a halo exchange.
The reference cannot process an image that does not fit in one process' memory at all.
"""
from __future__ import annotations

import os
import time
from concurrent.futures import ThreadPoolExecutor

from ..ops import host as H


def plan_strips(height: int, row_bytes: int, loops: int, mem_bytes: int, k: int | None = None,
                slots: int = 3) -> tuple[int, int]:
    """(rows per strip S, iterations per pass k) for a device-memory budget: every slot holds two
    ping-pong buffers of S + 2k rows."""
    k = max(1, min(int(k or 16), max(loops, 1)))
    per_row = 2 * slots * max(row_bytes, 1)
    rows = int(mem_bytes // per_row) - 2 * k
    s = max(rows, 2 * k, 16)
    return min(s, height), k


def run_file(path_in: str, path_out: str, width: int, height: int, channels: int, filt, loops: int,
             mem_bytes: int, k: int | None = None, device: int | None = None, rank: int = 0, world: int = 1,
             dist=None, io_threads: int = 2) -> dict:
    """Filter ``path_in`` ``loops`` times into ``path_out`` without ever holding more than ``mem_bytes``
    of it on the device. Returns a small report (passes, strips, seconds, bytes moved)."""
    from ..engine import ConvEngine
    rb = width * channels
    if H.file_size(path_in) < rb * height:
        raise ValueError(f"{path_in}: shorter than {width}x{height}x{channels}")
    S, kp = plan_strips(height, rb, loops, mem_bytes, k)
    nstrips = -(-height // S)
    npass = max(1, -(-loops // kp)) if loops > 0 else 0
    t0 = time.perf_counter()
    if loops <= 0:
        if rank == 0:
            data = H.read_rect(path_in, rb, 0, height, 0, rb)
            H.create_file(path_out, rb * height)
            H.write_rect(path_out, rb, 0, 0, data)
        return {"passes": 0, "strips": nstrips, "seconds": time.perf_counter() - t0, "strip_rows": S, "k": kp}
    tmp = path_out + ".pass"
    # arrange the ping-pong so that the LAST pass writes path_out
    outs = [(path_out if (npass - 1 - p) % 2 == 0 else tmp) for p in range(npass)]
    engines: dict[int, dict] = {}      # strip height (with halo) -> engine + its three staging slots
    pool = ThreadPoolExecutor(max_workers=max(1, io_threads))
    moved = 0

    def engine_for(rows: int) -> dict:
        if rows not in engines:
            eng = ConvEngine(width=width, height=rows, channels=channels, filter=filt, backend="fused",
                             device=device, slots=3, standalone=True)
            engines[rows] = {"eng": eng, "stage": [eng.pinned_buffers() for _ in range(3)], "n": 0,
                             "busy": [None, None, None]}     # per slot: (r0, r1, a0) of the step in flight
        return engines[rows]

    def retire(E: dict, slot: int, dst: str):
        """Wait for the step in flight on ``slot`` and hand its interior rows to a writer thread."""
        job = E["busy"][slot]
        if job is None:
            return None
        r0, r1, a0 = job
        E["eng"].slot_sync(slot)
        out = E["stage"][slot][1][0]
        E["busy"][slot] = None
        return pool.submit(H.write_rect, dst, rb, r0, 0, out[r0 - a0:r1 - a0])

    try:
        src, done = path_in, 0
        for p in range(npass):
            kk = min(kp, loops - done)
            dst = outs[p]
            if rank == 0:
                H.create_file(dst, rb * height)
            if dist is not None:
                dist.barrier()
            writes = []
            for i in range(rank, nstrips, world):
                r0, r1 = i * S, min(height, i * S + S)
                a0, a1 = max(0, r0 - kk), min(height, r1 + kk)
                E = engine_for(a1 - a0)
                slot = E["n"] % 3
                E["n"] += 1
                w = retire(E, slot, dst)                     # the slot's previous result leaves (async write)
                ins = E["stage"][slot][0][0]
                H.read_rect(src, rb, a0, a1 - a0, 0, rb, out=ins)    # overlaps the GPU work of the other slots
                if w is not None:
                    w.result()                               # its pinned output buffer is free again
                E["eng"].enqueue_step_pipelined(E["stage"][slot][0], E["stage"][slot][1], kk)
                E["busy"][slot] = (r0, r1, a0)
                moved += 2 * (a1 - a0) * rb
            for E in engines.values():
                for slot in range(3):
                    w = retire(E, slot, dst)
                    if w is not None:
                        writes.append(w)
            for w in writes:
                w.result()
            if dist is not None:
                dist.barrier()
            src, done = dst, done + kk
        if rank == 0 and npass > 1 and os.path.exists(tmp):
            os.remove(tmp)
    finally:
        pool.shutdown(wait=True)
        for E in engines.values():
            E["eng"].close()
    return {"passes": npass, "strips": nstrips, "strip_rows": S, "k": kp, "seconds": time.perf_counter() - t0,
            "h2d_d2h_bytes": moved, "ranks": world}
