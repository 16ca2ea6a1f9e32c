// Split-phase resident kernel with a PROTOCOL WARP. EXPERIMENTAL: written after the measurements of
// resident2_kernel.cuh; its flag protocol is replayed by scripts/sim_split_phase.py. First hardware
// contact at the very end of the round (scripts/r3_bringup.sh): exact on one rank (three filters, four
// depths) and on 2 / 4 / 9 virtual ranks, no hang — but not yet tuned: 0.45 ms per 100 iterations on
// a 1260x1920 RGB image at the planner's k=3 against 0.28 ms streaming (profiles/final_r1/
// modes_half_image.txt). Never selected automatically (mode "resident3" must be asked for); its GPU
// tests are opt-in (PCONV_TEST_RESIDENT3=1) because a named-barrier mistake would hang, not time out.
// Known next step: the protocol warp is a strictly sequential agent (poll late -> poll main -> wait for
// the strips -> fence + flag); releasing a tile's flag should not queue behind polling for the other
// tile — make it an event loop over non-blocking checks.
//
// resident2 showed that the halo loads themselves can land in the background, but every protocol step
// was bracketed by __syncthreads, so all 16 warps sat through the L2 round trips of the flag acquire
// and of the release fence (profiles/ncu_final/resident_R14_phase_stalls.txt). Here the 16th warp owns
// the protocol and the region is 15*R rows (a 17th warp would cost as much register file as four:
// warps are allocated in fours, 544 threads get 96 registers each instead of 128):
//   * compute warps (0..14) only ever meet each other, at the named barrier kBarCompute (480 threads);
//   * after storing a tile's strips they ARRIVE (non-blocking) on kBarOut+slot; the protocol warp SYNCs on
//     it, then fences and releases the tile's epoch flag (+ ring mirrors) — the fence stalls nobody else;
//   * the protocol warp polls the neighbour flags a tile needs next (one lane per direction) and ARRIVES
//     on kBarInMain+slot / kBarInLate; compute warps SYNC on it right before they issue the halo loads —
//     with the split-phase slack the flags are normally there already, so that sync is free.
// Schedule, wait sets, epochs and buffers are exactly resident2's. Barrier generations cannot overlap:
// between two arrivals on one id the protocol warp passes a kBarOut sync, which needs the compute
// warps to have passed the earlier generation. Ordering: neighbour's stores -> its release -> protocol
// lane's acquire -> barrier -> compute warps' loads (causality order is transitive through barriers).
#pragma once

#include <cuda_runtime.h>

#include "pconv_cuda.h"
#include "region_io.cuh"
#include "resident2_kernel.cuh"
#include "resident_kernel.cuh"
#include "runtime_internal.h"
#include "stencil_core.cuh"

namespace pconv {

constexpr int kCompWarps = kWarps - 1;                 // warps 0..14 compute, warp 15 runs the protocol
constexpr int kCompThreads = 32 * kCompWarps;
constexpr int kProtoThreads = kThreads;                // everybody (named-barrier hand-offs)
constexpr int kBarCompute = 1, kBarOut = 2 /* + slot */, kBarInMain = 4 /* + slot */, kBarInLate = 6;

__device__ __forceinline__ void bar_sync(int id, int count) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
}
__device__ __forceinline__ void bar_arrive(int id, int count) {
    asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory");
}

// region_iterate (stencil_core.cuh) with the compute-warp barrier instead of __syncthreads
template <int KIND, int C, int R, bool MASKED>
__device__ __forceinline__ void region_iterate_named(uint4 (&rows)[R], EdgeBuffers& eb, int k,
                                                     const uint4& cmask, uint32_t rowmask, bool need_cols,
                                                     const FilterDev& f) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint4 zero = make_uint4(0u, 0u, 0u, 0u);
    for (int it = 0; it < k; ++it) {
        const int par = it & 1;
        eb.row[par][warp][0][lane] = rows[0];
        eb.row[par][warp][1][lane] = rows[R - 1];
        bar_sync(kBarCompute, kCompThreads);
        const uint4 top = warp > 0 ? eb.row[par][warp - 1][1][lane] : zero;
        const uint4 bottom = warp < kCompWarps - 1 ? eb.row[par][warp + 1][0][lane] : zero;
        band_iterate<KIND, C, R, MASKED>(rows, top, bottom, cmask, rowmask, need_cols, f);
    }
}

template <int KIND, int C, int R>
__global__ void __launch_bounds__(kThreads, 1)
resident3_kernel(const __grid_constant__ ResidentArgs A) {
    __shared__ EdgeBuffers eb;
    __shared__ unsigned long long s_wait_ns;
    const pconv_geometry_t& g = A.g;
    constexpr int RH = kCompWarps * R;                              // region rows (15 compute warps)
    const int k = A.k;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const bool is_proto = warp == kCompWarps;                       // warp 15
    const int tip = blockIdx.x / A.ntx, tj = blockIdx.x % A.ntx;   // physical CTA row / column
    const int64_t gb0 = g.cb_lo[1] + (int64_t)tj * A.TWB - A.HB;
    const int64_t gb = gb0 + 16 * lane;
    const int fw = A.ntx + 2;  // width of the extended flag grid
    const int64_t tile_rows = g.row_lo[2] - g.row_lo[1];

    if (threadIdx.x == 0) s_wait_ns = 0ull;
    __syncthreads();

    // ---- launch-level protocol: the neighbours' input tiles are in place ----------------------------
    const int64_t gr_top = g.row_lo[1] + (int64_t)(2 * tip) * A.TH - k;          // both regions together
    const int64_t gr_end = g.row_lo[1] + (int64_t)(2 * tip + 1) * A.TH - k + RH;
    if (A.sync.wait_value != 0u && threadIdx.x < 9 && threadIdx.x != 4) {
        const int dy = (int)threadIdx.x / 3, dx = (int)threadIdx.x % 3;
        const bool rows_hit = dy == 0 ? gr_top < g.row_lo[1] : dy == 2 ? gr_end > g.row_lo[2] : true;
        const bool cols_hit = dx == 0 ? gb0 < g.cb_lo[1] : dx == 2 ? gb0 + kSP > g.cb_lo[2] : true;
        if (rows_hit && cols_hit && A.src[0].p[threadIdx.x] != nullptr) {
            const uint64_t ns = wait_flag(A.sync.local_flags + threadIdx.x, A.sync.wait_value,
                                          A.sync.timeout_ns, A.sync.status);
            atomicMax(&s_wait_ns, (unsigned long long)ns);
        }
    }
    __syncthreads();

    // ---- the two virtual tiles (identical to resident2) -------------------------------------------------
    bool need_cols;
    const uint4 cmask = column_mask(g, gb0, &need_cols);
    const int vcb = (int)min((int64_t)A.TWB, g.cb_lo[2] - g.cb_lo[1] - (int64_t)tj * A.TWB);
    const bool col_int = (16 * lane >= A.HB) && (16 * lane < A.HB + vcb);
    const bool col_full = (16 * lane >= A.HB) && (16 * lane + 16 <= A.HB + vcb);
    const bool col_edge = col_int && (16 * lane < 2 * A.HB || 16 * lane + 16 > vcb);

    auto describe = [&](int s, bool first) {
        VTile t;
        t.vi = 2 * tip + s;
        const int64_t gr0 = g.row_lo[1] + (int64_t)t.vi * A.TH - k;
        t.grw = gr0 + (int64_t)warp * R;
        const int vr = (int)max((int64_t)0, min((int64_t)A.TH, tile_rows - (int64_t)t.vi * A.TH));
        const int sd = s == 0 ? 1 : -1;            // where the sibling row lies: +1 below, -1 above
        t.m_publish = t.m_main = t.m_late = t.m_final = t.rowmask = 0u;
#pragma unroll
        for (int i = 0; i < R; ++i) {
            const int rr = warp * R + i;
            const bool row_int = rr >= k && rr < k + vr;
            const bool row_edge = rr < 2 * k || rr >= vr;
            const bool sib_side = sd > 0 ? rr >= k + vr : rr < k;
            if (row_int && col_int && (row_edge || col_edge)) t.m_publish |= 1u << i;
            if (!(row_int && col_full)) {
                if (first && sib_side) t.m_late |= 1u << i;
                else t.m_main |= 1u << i;
            }
            if (row_int && col_int) t.m_final |= 1u << i;
            if (t.grw + i >= 0 && t.grw + i < g.img_rows) t.rowmask |= 1u << i;
        }
        t.masked = need_cols || gr0 < 0 || gr0 + RH > g.img_rows;
        t.first = first;
        const uint32_t row_up = 0x007u, row_mid = 0x028u, row_dn = 0x1c0u;   // d = 0..2 | 3,5 | 6..8
        const uint32_t sib_row = sd > 0 ? row_dn : row_up, far_row = sd > 0 ? row_up : row_dn;
        const uint32_t sib_self = sd > 0 ? 0x080u : 0x002u;
        if (first) { t.w_main = far_row | row_mid; t.w_late = sib_row & ~sib_self; }
        else       { t.w_main = (far_row | row_mid | sib_row) & ~sib_self; t.w_late = 0u; }
        return t;
    };
    const int s_first = tip & 1;
    VTile tc = describe(s_first, true), to = describe(1 - s_first, false);
    const int nblocks = (A.loops + k - 1) / k;

    if (is_proto) {
        // =================================== protocol warp ==============================================
        // lane d (< 9, != 4) polls the flag of neighbour direction d of virtual tile row vi
        auto poll = [&](int vi, uint32_t wset, uint32_t epoch) {
            if (lane < 9 && ((wset >> lane) & 1u)) {
                const int ei = vi + lane / 3, ej = tj + lane % 3;       // extended-grid coordinates
                const int sy = ei < 1 ? 0 : (ei > A.nty ? 2 : 1), sx = ej < 1 ? 0 : (ej > A.ntx ? 2 : 1);
                const bool exists = (sy == 1 && sx == 1) || A.src[0].p[sy * 3 + sx] != nullptr;
                if (exists) {
                    const uint32_t* fl = A.cta_flags + ei * fw + ej;
                    const uint64_t ns = (sy == 1 && sx == 1)
                        ? wait_flag<false>(fl, epoch, A.sync.timeout_ns, A.sync.status)
                        : wait_flag<true>(fl, epoch, A.sync.timeout_ns, A.sync.status);
                    atomicMax(&s_wait_ns, (unsigned long long)ns);
                }
            }
            __syncwarp();
        };
        for (int hp = 0; hp < 2 * nblocks; ++hp) {
            const int b = hp >> 1;
            if (b > 0 && tc.first) {                      // zero-slack part: the sibling row's laterals
                poll(tc.vi, tc.w_late, A.epoch_base + (uint32_t)b);
                bar_arrive(kBarInLate, kProtoThreads);
            }
            const int bo = (hp + 1) >> 1;
            if (bo > 0 && bo < nblocks) {                 // what the idle tile needs for its next block
                poll(to.vi, to.w_main, A.epoch_base + (uint32_t)bo);
                bar_arrive(kBarInMain + (to.first ? 0 : 1), kProtoThreads);
            }
            if (b < nblocks - 1) {                        // the strips of `cur` are stored: publish its epoch
                const uint32_t epoch = A.epoch_base + (uint32_t)(b + 1);
                bar_sync(kBarOut + (tc.first ? 0 : 1), kProtoThreads);
                if (lane == 0)
                    st_release_gpu(A.cta_flags + (tc.vi + 1) * fw + (tj + 1), epoch);
                if (lane < 9 && lane != 4) {
                    const int sy = lane / 3 - 1, sx = lane % 3 - 1;
                    const bool on_y = sy == 0 || (sy < 0 ? tc.vi == 0 : tc.vi == A.nty - 1);
                    const bool on_x = sx == 0 || (sx < 0 ? tj == 0 : tj == A.ntx - 1);
                    uint32_t* remote = A.remote_cta_flags[lane];
                    if (on_y && on_x && remote != nullptr)
                        st_release_sys(remote + (tc.vi + 1 - sy * A.nty) * fw + (tj + 1 - sx * A.ntx), epoch);
                }
                __syncwarp();
            }
            const VTile t = tc; tc = to; to = t;
        }
    } else {
        // =================================== compute warps ==============================================
        uint4 cur[R], oth[R];
        load_band<R>(cur, g, A.src[0], tc.grw, gb);
        load_band<R>(oth, g, A.src[0], to.grw, gb);
        for (int hp = 0; hp < 2 * nblocks; ++hp) {
            const int b = hp >> 1;                               // block computed on `cur` now
            const int kb = min(k, A.loops - b * k);
            if (b > 0 && tc.first) {
                // the sibling's strips were stored before the kBarCompute sync of its publish; the
                // lateral corner owners' flags have been seen by the protocol warp
                bar_sync(kBarInLate, kProtoThreads);
                load_band<R, true>(cur, g, A.src[b & 1], tc.grw, gb, tc.m_late);
            }
            for (int part = 0; part < 2; ++part) {
                const int kk = part == 0 ? kb / 2 : kb - kb / 2;
                if (tc.masked) region_iterate_named<KIND, C, R, true>(cur, eb, kk, cmask, tc.rowmask, need_cols, A.f);
                else region_iterate_named<KIND, C, R, false>(cur, eb, kk, cmask, tc.rowmask, need_cols, A.f);
                bar_sync(kBarCompute, kCompThreads);   // edge buffers: the next call restarts at parity 0
                if (part == 0) {
                    const int bo = (hp + 1) >> 1;
                    if (bo > 0 && bo < nblocks) {
                        bar_sync(kBarInMain + (to.first ? 0 : 1), kProtoThreads);
                        load_band<R, true>(oth, g, A.src[bo & 1], to.grw, gb, to.m_main);
                    }
                }
            }
            if (b < nblocks - 1) {
                store_band<R>(cur, tc.m_publish, A.self[(b + 1) & 1], g, tc.grw, gb);
                bar_sync(kBarCompute, kCompThreads);        // every strip is stored (sibling reads them next)
                bar_arrive(kBarOut + (tc.first ? 0 : 1), kProtoThreads);
            } else {
                store_band<R>(cur, tc.m_final, A.self[nblocks & 1], g, tc.grw, gb);
            }
            swap_bands<R>(cur, oth, tc, to);
        }
    }
    __syncthreads();
    if (threadIdx.x == 0 && s_wait_ns) record_wait(A.sync.status, s_wait_ns);

    // ---- launch-level protocol: publish "run done" -------------------------------------------------
    if (A.sync.signal_value != 0u && threadIdx.x == 0) {   // acq_rel ticket: releases this CTA's stores
        const uint32_t ticket = atom_add_acq_rel_gpu(A.sync.done_counter, 1u);
        if (ticket == gridDim.x - 1) {
            *A.sync.done_counter = 0u;
            if (!A.sync.inject_drop_signal) {
                fence_acq_rel_sys();
#pragma unroll
                for (int d = 0; d < 9; ++d)
                    if (d != 4 && A.sync.remote_flags[d] != nullptr)
                        st_release_sys(A.sync.remote_flags[d], A.sync.signal_value);
            }
        }
    }
}

#define PCONV_RES3_R_SEP(X) X(2) X(4) X(5) X(7)
#define PCONV_RES3_R_INT(X) X(2) X(4) X(5)
#define PCONV_RES3_R_F32(X) X(2) X(4) X(5)

template <int KIND, int C>
static ResidentFn pick_resident3_r(int R) {
#define CASE(RR) if (R == RR) return resident3_kernel<KIND, C, RR>;
    if constexpr (KIND == PCONV_KIND_F32) { PCONV_RES3_R_F32(CASE) }
    else if constexpr (KIND == PCONV_KIND_INT) { PCONV_RES3_R_INT(CASE) }
    else { PCONV_RES3_R_SEP(CASE) }
#undef CASE
    return nullptr;
}

template <int KIND>
static ResidentFn pick_resident3_kind(int C, int R) {
    switch (C) {
        case 1: return pick_resident3_r<KIND, 1>(R);
        case 3: return pick_resident3_r<KIND, 3>(R);
        case 4: return pick_resident3_r<KIND, 4>(R);
        default: return nullptr;
    }
}

ResidentFn pick_resident3_f32(int C, int R);
ResidentFn pick_resident3_int(int C, int R);
ResidentFn pick_resident3_sep(int C, int R);

}  // namespace pconv
