#!/usr/bin/env python
"""Discrete simulation of the split-phase resident kernel's flag protocol (resident2_kernel.cuh): every
CTA runs its program (waits on per-tile epoch flags, publishes), in an arbitrary fair interleaving; the
script reports a deadlock (no CTA can advance) with the blocked waits. Development aid — no GPU needed.

    python scripts/sim_split_phase.py GRID_ROWS GRID_COLS NTY NTX NBLOCKS
"""
import sys


def programs(nty, ntx, nblocks):
    """Per physical CTA (tip, tj): list of ops ('wait', vi, dirs, epoch) / ('pub', vi, epoch)."""
    ROW_UP, ROW_MID, ROW_DN = [0, 1, 2], [3, 5], [6, 7, 8]
    progs = {}
    for tip in range(nty // 2):
        for tj in range(ntx):
            def tile(s, first):
                sd = 1 if s == 0 else -1
                sib_row, far_row = (ROW_DN, ROW_UP) if sd > 0 else (ROW_UP, ROW_DN)
                sib_self = 7 if sd > 0 else 1
                if first:
                    return dict(vi=2 * tip + s, first=True, w_main=far_row + ROW_MID,
                                w_late=[d for d in sib_row if d != sib_self])
                return dict(vi=2 * tip + s, first=False,
                            w_main=[d for d in far_row + ROW_MID + sib_row if d != sib_self], w_late=[])
            s_first = tip & 1
            tc, to = tile(s_first, True), tile(1 - s_first, False)
            ops = []
            for hp in range(2 * nblocks):
                b = hp >> 1
                if b > 0 and tc["first"]:
                    ops.append(("wait", tc["vi"], tc["w_late"], b))
                bo = (hp + 1) >> 1
                if 0 < bo < nblocks:
                    ops.append(("wait", to["vi"], to["w_main"], bo))
                if b < nblocks - 1:
                    ops.append(("pub", tc["vi"], b + 1))
                tc, to = to, tc
            progs[(tip, tj)] = ops
    return progs


def simulate(GR, GC, nty, ntx, nblocks):
    """Run every CTA of a GR x GC rank grid to completion; returns the blocked (rank, cta, op) list."""
    fw = ntx + 2
    flags = {(r, c): [0] * ((nty + 2) * fw) for r in range(GR) for c in range(GC)}
    progs = {(r, c): programs(nty, ntx, nblocks) for r in range(GR) for c in range(GC)}
    pc = {(rk, cta): 0 for rk in progs for cta in progs[rk]}

    def exists(rk, sy, sx):
        r, c = rk[0] + sy - 1, rk[1] + sx - 1
        return 0 <= r < GR and 0 <= c < GC

    def ready(rk, tj, vi, dirs, epoch):
        for d in dirs:
            ei, ej = vi + d // 3, tj + d % 3
            sy = 0 if ei < 1 else (2 if ei > nty else 1)
            sx = 0 if ej < 1 else (2 if ej > ntx else 1)
            if not ((sy == 1 and sx == 1) or exists(rk, sy, sx)):
                continue
            if flags[rk][ei * fw + ej] < epoch:
                return False, (d, ei, ej)
        return True, None

    def publish(rk, tj, vi, epoch):
        flags[rk][(vi + 1) * fw + tj + 1] = epoch
        for t in range(9):
            if t == 4:
                continue
            sy, sx = t // 3 - 1, t % 3 - 1
            on_y = sy == 0 or (vi == 0 if sy < 0 else vi == nty - 1)
            on_x = sx == 0 or (tj == 0 if sx < 0 else tj == ntx - 1)
            nb = (rk[0] + sy, rk[1] + sx)
            if on_y and on_x and nb in flags and nb != rk:
                flags[nb][(vi + 1 - sy * nty) * fw + (tj + 1 - sx * ntx)] = epoch

    progress = True
    while progress:
        progress = False
        for (rk, cta), i in list(pc.items()):
            ops = progs[rk][cta]
            while i < len(ops):
                op = ops[i]
                if op[0] == "wait":
                    ok, _ = ready(rk, cta[1], op[1], op[2], op[3])
                    if not ok:
                        break
                else:
                    publish(rk, cta[1], op[1], op[2])
                i += 1
                progress = True
            pc[(rk, cta)] = i
    return [(rk, cta, progs[rk][cta][i], ready(rk, cta[1], *progs[rk][cta][i][1:4])[1])
            for (rk, cta), i in pc.items() if i < len(progs[rk][cta])]


def main():
    GR, GC, nty, ntx, nblocks = (int(v) for v in sys.argv[1:6])
    stuck = simulate(GR, GC, nty, ntx, nblocks)
    if not stuck:
        print("OK: every CTA ran to completion")
        return 0
    print(f"DEADLOCK: {len(stuck)} CTAs blocked")
    for rk, cta, op, why in stuck[:12]:
        print(f"  rank {rk} cta {cta}: {op} blocked on dir/ei/ej {why}")
    return 1


if __name__ == "__main__":
    sys.exit(main())
