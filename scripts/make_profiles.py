#!/usr/bin/env python
"""Regenerate the committed evidence under profiles/ from the build tree and gpurun_out/:

  profiles/sass_<kernel>.txt      hot-loop SASS of the flagship kernels + opcode histogram + the
                                  system-scope flag / peer-load instructions of the halo protocol
  profiles/ptxas_resources.txt    registers / spills / shared memory of every compiled kernel
  profiles/ncu_<name>.csv/.md     key metrics of each gpurun_out/*.ncu-rep (raw page, filtered)
"""
import csv
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "parallel_convolution_b200", "_lib")
OUT = os.path.join(ROOT, "profiles")
os.makedirs(OUT, exist_ok=True)

KERNELS = {
    "stream_sep121_rgb_R14": ("stream_kind_sep.cu.o", "_ZN5pconv13stream_kernelILi2ELi3ELi14EEEvNS_10StreamArgsE"),
    "stream_sep121_grey_R14": ("stream_kind_sep.cu.o", "_ZN5pconv13stream_kernelILi2ELi1ELi14EEEvNS_10StreamArgsE"),
    "stream_int_rgb_R10": ("stream_kind_int.cu.o", "_ZN5pconv13stream_kernelILi1ELi3ELi10EEEvNS_10StreamArgsE"),
    "stream_f32_grey_R8": ("stream_kind_f32.cu.o", "_ZN5pconv13stream_kernelILi0ELi1ELi8EEEvNS_10StreamArgsE"),
    "persist_sep121_rgb_R14_W16": ("persist_sep_w16.cu.o", "_ZN5pconv14persist_kernelILi2ELi3ELi14ELi16EEEvNS_11PersistArgsE"),
    "persist_sep121_rgb_R7_W8": ("persist_sep_w8.cu.o", "_ZN5pconv14persist_kernelILi2ELi3ELi7ELi8EEEvNS_11PersistArgsE"),
    "persist_int_rgb_R10_W16": ("persist_int_w16.cu.o", "_ZN5pconv14persist_kernelILi1ELi3ELi10ELi16EEEvNS_11PersistArgsE"),
}


def sass(obj, fn):
    out = subprocess.run(["cuobjdump", "-sass", "-fun", fn, os.path.join(LIB, obj)], capture_output=True, text=True).stdout
    lines = [re.sub(r"/\* 0x[0-9a-f]+ \*/", "", l).rstrip() for l in out.splitlines() if re.match(r"\s+/\*[0-9a-f]{4,5}\*/", l)]
    return lines


def opcode(line):
    body = re.sub(r"/\*[0-9a-f]+\*/", "", line).strip()
    body = re.sub(r"^@!?U?P[0-9T]+\s+", "", body)
    return body.split()[0].rstrip(";") if body else ""


def segment_report(rep_path, out_path):
    """Warp-stall samples of a persistent-kernel capture, aggregated between its synchronisation points
    (barriers, fences, bulk copy, mbarrier): where the time of an exchange goes, from the ncu source page."""
    src = subprocess.run(["ncu", "-i", rep_path, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(src.splitlines()))
    if len(rows) < 4:
        return
    hdr, data = rows[1], rows[2:]
    ci = {n: i for i, n in enumerate(hdr)}
    if "# Samples" not in ci:
        return
    S, SRC, IE = ci["# Samples"], ci["Source"], ci["Instructions Executed"]
    stalls = [c for c in hdr if c.startswith("stall_") and "Not Issued" not in c]
    tot = sum(int(r[S]) for r in data) or 1
    segs, acc, start = [], 0, 0
    for i, r in enumerate(data):
        acc += int(r[S])
        if any(m in r[SRC] for m in ("BAR.SYNC", "UBLKCP", "SYNCS.ARRIVE", "MEMBAR.ALL.GPU", "MEMBAR.ALL.SYS", "NANOSLEEP")):
            segs.append((start, i + 1, acc, r[SRC].strip()[:44]))
            acc, start = 0, i + 1
    segs.append((start, len(data), acc, "end"))
    with open(out_path, "w") as f:
        f.write(f"# {rows[0][1] if len(rows[0]) > 1 else ''}\n# {os.path.basename(rep_path)}: {tot} warp-stall samples over {len(data)} SASS instructions\n")
        f.write("# segment = instructions up to and including the named synchronisation instruction\n")
        for a, b, n, label in segs:
            if n < tot * 0.003:
                continue
            st = {c: sum(int(data[j][ci[c]]) for j in range(a, b)) for c in stalls}
            top = sorted(st.items(), key=lambda kv: -kv[1])[:3]
            ie = sum(int(data[j][IE]) for j in range(a, b))
            f.write(f"instr[{a:5d},{b:5d})  samples {n:6d} ({100 * n / tot:5.1f} %)  warp-instructions {ie:10d}  ends at {label:44s} "
                    f"top stalls: {', '.join(k.replace('stall_', '') + '=' + str(v) for k, v in top)}\n")
        worst = sorted(range(len(data)), key=lambda j: -int(data[j][S]))[:12]
        f.write("# hottest instructions\n")
        for j in sorted(worst):
            f.write(f"  [{j:5d}] {int(data[j][S]):6d}  {data[j][SRC].strip()[:90]}\n")
    print("wrote", out_path)


def main():
    for name, (obj, fn) in KERNELS.items():
        L = sass(obj, fn)
        if not L:
            continue
        bars = [i for i, l in enumerate(L) if "BAR.SYNC" in l]
        shfl = [i for i, l in enumerate(L) if "SHFL" in l]
        # the band iteration loops are the regions between barriers that contain shuffles; keep the shortest
        # complete one (the unmasked variant)
        loops = []
        for a, b in zip(bars, bars[1:]):
            n = sum(1 for i in shfl if a < i < b)
            if n >= 4:
                loops.append((b - a, a, b))
        hist = {}
        for l in L:
            op = opcode(l).split(".")[0]
            hist[op] = hist.get(op, 0) + 1
        with open(os.path.join(OUT, f"sass_{name}.txt"), "w") as f:
            f.write(f"# {fn}\n# cuobjdump -sass of parallel_convolution_b200/_lib/{obj} (sm_100a), {len(L)} instructions\n")
            f.write("# whole-kernel opcode histogram: " + ", ".join(f"{k}:{v}" for k, v in sorted(hist.items(), key=lambda kv: -kv[1])[:14]) + "\n")
            proto = [l for l in L if re.search(r"STRONG\.SYS|\.SYS|MEMBAR|ATOM|CCTL|NANOSLEEP|GLOBALTIMER|ERRBAR|FENCE|UBLKCP|SYNCS|UTMA", l)]
            f.write("#\n# --- halo protocol instructions: system/gpu-scope flags and fences, bulk async copy (UBLKCP = "
                    "cp.async.bulk), mbarrier (SYNCS.*) ---\n")
            f.write("\n".join(proto[:80]) + "\n")
            ldst = [l for l in L if re.search(r"LDG\.E\.128|STG\.E\.128", l)]
            f.write(f"#\n# --- 128-bit global accesses (tile / peer-tile rows straight to/from registers): {len(ldst)} sites ---\n")
            f.write("\n".join(ldst[:12]) + "\n")
            if loops:
                n, a, b = min(loops)
                body = L[a:b + 1]
                h2 = {}
                for l in body:
                    op = opcode(l).split(".")[0]
                    h2[op] = h2.get(op, 0) + 1
                f.write(f"#\n# --- one fused iteration over the register band (barrier to barrier, {n} instructions) ---\n")
                f.write("# opcode histogram: " + ", ".join(f"{k}:{v}" for k, v in sorted(h2.items(), key=lambda kv: -kv[1])) + "\n")
                f.write("\n".join(body) + "\n")
        print("wrote", f"profiles/sass_{name}.txt", len(L), "instructions")

    # ptxas resources
    rows = []
    for log in sorted(os.listdir(LIB)):
        if not log.endswith(".ptxas.log"):
            continue
        cur = None
        for line in open(os.path.join(LIB, log)):
            m = re.search(r"Compiling entry function '(\S+)'", line)
            if m:
                cur = m.group(1)
            m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", line)
            if m:
                sp = m.groups()
            m = re.search(r"Used (\d+) registers, used (\d+) barriers(?:, (\d+) bytes cumulative stack size)?(?:, (\d+) bytes smem)?", line)
            if m and cur:
                rows.append((cur, m.group(1), sp[1], sp[2], m.group(4) or "0"))
                cur = None
    with open(os.path.join(OUT, "ptxas_resources.txt"), "w") as f:
        f.write("# nvcc -gencode arch=compute_100a,code=sm_100a -Xptxas -v : registers / spill bytes / static smem per kernel\n")
        f.write("# template args: <KIND (0 f32, 1 int, 2 sep121), CHANNEL STRIDE, BAND HEIGHT R>\n")
        for r in rows:
            name = subprocess.run(["c++filt", r[0]], capture_output=True, text=True).stdout.strip() or r[0]
            name = re.sub(r"\(.*", "", name)
            f.write(f"{name:70s} regs={r[1]:>3s} spill_st={r[2]:>3s} spill_ld={r[3]:>3s} smem={r[4]}\n")
    print("wrote profiles/ptxas_resources.txt", len(rows), "kernels")

    # ncu reports
    want = ["gpu__time_duration.sum", "sm__cycles_elapsed.avg.per_second", "sm__inst_executed.avg.per_cycle_active",
            "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
            "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tc", "sm__warps_active.avg.pct_of_peak_sustained_active",
            "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
            "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "smsp__inst_executed.sum",
            "smsp__average_warps_issue_stalled", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_bytes.sum",
            "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput", "lts__throughput"]
    gdir = os.path.join(ROOT, "gpurun_out", sys.argv[1] if len(sys.argv) > 1 else "final_r2")
    for rep in sorted(os.listdir(gdir)) if os.path.isdir(gdir) else []:
        if not rep.endswith(".ncu-rep"):
            continue
        raw = subprocess.run(["ncu", "-i", os.path.join(gdir, rep), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        segment_report(os.path.join(gdir, rep), os.path.join(OUT, f"ncu_{rep[:-8]}_segments.txt"))
        rows = list(csv.reader(raw.splitlines()))
        if len(rows) < 3:
            continue
        hdr, units = rows[0], rows[1]
        base = rep[:-8]
        with open(os.path.join(OUT, f"ncu_{base}.csv"), "w") as f:
            f.write("kernel,metric,unit,value\n")
            for vals in rows[2:]:
                kname = vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
                for h, u, v in zip(hdr, units, vals):
                    if any(h.startswith(w) or w in h for w in want):
                        f.write(f"{kname},{h},{u},{v}\n")
        print("wrote", f"profiles/ncu_{base}.csv")


if __name__ == "__main__":
    sys.exit(main())
