// Launch planner (native): region geometry and temporal-blocking depth for the sm_100a kernels.
//
// A CTA region is RH = 16*R rows x 512 bytes held in registers (R = band height per warp), of which
// the interior TH x TWB (TH <= RH - 2k, TWB = 512 - 2*HB, HB = roundup16(k*C)) is written back
// after k fused iterations. Deeper k divides the number of launches / cross-GPU synchronisations
// by k (every BASELINE config is latency- or ALU-bound, never link-bound: SURVEY.md §7.2) but
// recomputes a thicker halo; the planner minimises a simple time model over (k, R, TH) including
// the wave quantisation over the SMs. The reference has no scheduler at all: one launch / one MPI
// exchange per iteration, block size hard-coded to 16 (cuda/cuda_convolute.cu:64).
//
// The plan must be identical on every rank, so it only depends on quantities every rank knows.
#include <algorithm>
#include <cmath>
#include <cstdint>

#include "pconv_host.h"

namespace {

constexpr int kRegionBytes = 512, kWarps = 16;
// Modelled SM throughput in region bytes per clock (B200 measurements, profiles/): only used to
// rank candidate plans, never reported as a result.
double bytes_per_clk(int kind) {
    return kind == PCONV_KIND_SEP121 ? 23.0 : (kind == PCONV_KIND_INT ? 9.0 : 4.0);
}
constexpr double kClockGHz = 1.9, kLaunchUs = 6.0, kLoadBytesPerUsPerSm = 40e3;
// Persistent kernel (persist_kernel.cuh): what one exchange costs a CTA — push (~2.3 us: the LSU cost is
// per store instruction), release + flag round trip + bulk fetch (~2.6 us of which ~1.2 us is waiting for
// the slowest neighbour), inbox -> registers (~1 us). A pushed exchange costs about the same across NVLink
// (posted stores, local polling). Two CTAs sharing an SM do NOT hide each other's exchanges — every CTA of
// the grid is ready at the same moment — they roughly halve the per-CTA cost instead. Measured on B200
// (profiles/README.md); only used to rank plans — the autotuner measures the best few.
constexpr double kExchangeUs = 6.3, kExchangeUsPeer = 7.5, kExchangeUsHidden = 3.7, kExchangeUsHiddenPeer = 4.5;

int roundup(int v, int m) { return (v + m - 1) / m * m; }
int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

}  // namespace

extern "C" int pconv_plan(const pconv_plan_request_t* q, pconv_plan_t* out) {
    if (!q || !out || q->nbands < 1 || q->tile_rows < 1 || q->tile_cb < 1 || q->C < 1) return -1;
    const int loops = std::max(q->loops, 1);
    const int sms = q->sms > 0 ? q->sms : 148;
    const double rate = bytes_per_clk(q->kind) * kClockGHz * 1e3;  // region bytes per us per SM
    int max_band = 0;
    for (int i = 0; i < q->nbands; ++i) max_band = std::max(max_band, q->bands[i]);
    const int warps = q->resident ? (q->warps == 8 ? 8 : kWarps) : kWarps;   // persistent CTAs: 8 or 16 warps
    const int per_sm_cap = kWarps / warps;                                     // co-resident CTAs per SM
    int kmax = std::min({loops, (warps * max_band - 1) / 2, (kRegionBytes / 2 - 64) / q->C});   // (upper bound)
    if (q->min_nbr_rows > 0) kmax = (int)std::min<int64_t>(kmax, q->min_nbr_rows);
    if (q->min_nbr_cols > 0) kmax = (int)std::min<int64_t>(kmax, q->min_nbr_cols);
    kmax = std::max(kmax, 1);
    const int k_lo = q->force_k > 0 ? q->force_k : 1, k_hi = q->force_k > 0 ? q->force_k : kmax;
    bool found = false;
    double best = 0.0;
    for (int k = k_lo; k <= k_hi; ++k) {
        const int HB = roundup(std::max(k * q->C, 1), 16), TWB = kRegionBytes - 2 * HB;
        if (TWB < std::max(64, q->resident ? HB : 0)) continue;
        const int64_t ntx = ceil_div(q->tile_cb, TWB);
        for (int bi = 0; bi < q->nbands; ++bi) {
            const int R = q->bands[bi];
            if (q->force_R > 0 && R != q->force_R) continue;
            const int RH = warps * R, th_max = RH - 2 * k;
            if (th_max < (q->resident ? k : 1)) continue;
            const int64_t nty = ceil_div(q->tile_rows, th_max);
            if (q->resident && ntx * nty > (int64_t)per_sm_cap * sms) continue;   // every CTA co-resident
            if (q->resident && ntx * nty > 1024) continue;                         // flag area (xplan.h)
            int TH = (int)ceil_div(q->tile_rows, nty);      // balance rows over the same tile count
            if (q->resident) TH = std::max(TH, k);
            if (q->resident) {
                // A CTA exchanges with its eight surrounding CTAs only: with a rank below / to the right,
                // the last CTA row / column must itself be at least one halo deep, or the halo of the row
                // / column before it would reach past it into that rank; across ranks whole chunks move.
                if (ceil_div(q->tile_rows, TH) != nty) continue;
                const int64_t last_rows = q->tile_rows - (nty - 1) * (int64_t)TH;
                const int64_t last_cb = q->tile_cb - (ntx - 1) * (int64_t)TWB;
                if (q->min_nbr_rows > 0 && last_rows < k) continue;
                if (q->min_nbr_cols > 0 && (last_cb < HB || (q->tile_cb & 15))) continue;
            }
            const int64_t waves = ceil_div(ntx * nty, sms);
            double total, per_launch;
            if (q->resident) {
                const int64_t nblocks = ceil_div(loops, k);
                const double on_sm = (double)ceil_div(ntx * nty, sms);       // CTAs sharing the busiest SM
                const double t_x = on_sm > 1.0 ? (q->peers ? kExchangeUsHiddenPeer : kExchangeUsHidden)
                                               : (q->peers ? kExchangeUsPeer : kExchangeUs);
                // a lone 8-warp CTA cannot fill the SM's issue slots (measured: ~15 % slower than 16 warps)
                const double r_eff = (warps == 8 && on_sm == 1.0) ? 0.85 * rate : rate;
                total = kLaunchUs + on_sm * (RH * (double)kRegionBytes + (double)TH * TWB) / kLoadBytesPerUsPerSm +
                        (double)loops * on_sm * (RH * (double)kRegionBytes / r_eff) + (nblocks - 1) * on_sm * t_x;
                per_launch = total;
            } else {
                const double t_tile = RH * (double)kRegionBytes / kLoadBytesPerUsPerSm +
                                      k * RH * (double)kRegionBytes / rate +
                                      (double)TH * TWB / kLoadBytesPerUsPerSm;
                per_launch = kLaunchUs + waves * t_tile;
                total = ceil_div(loops, k) * per_launch;
            }
            if (!found || total < best - 1e-9) {
                found = true;
                best = total;
                out->k = k; out->R = R; out->TH = TH; out->TWB = TWB; out->HB = HB;
                out->ntx = (int32_t)ntx; out->nty = (int32_t)nty; out->waves = (int32_t)waves;
                out->est_us = per_launch;
            }
        }
    }
    return found ? 0 : 1;   // 1: no feasible plan (resident: the tile does not fit on chip)
}
