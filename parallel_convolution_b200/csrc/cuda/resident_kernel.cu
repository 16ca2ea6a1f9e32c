// Resident kernel: one persistent launch per run. The image tile is spread over the CTAs of the
// GPU (one region per CTA, held in registers — stencil_core.cuh) and stays on chip for ALL
// iterations; only k-deep border strips move, every k iterations:
//
//   exchange b (b = 1, 2, ...):
//     1. each CTA stores the border strips of its interior into tile buffer parity (b & 1)
//     2. st.release (gpu scope) of epoch_base+b to its own flag — and, for CTAs
//        on the edge of the GPU's tile, to the mirrored ring entry in the neighbour GPU's flag
//        array (peer store over NVLink)
//     3. eight threads spin (ld.acquire.sys) on the flags of the eight surrounding CTAs — local
//        CTAs or ring entries written by the neighbour GPU
//     4. the halo rows / columns of the region are re-loaded from the surrounding owners' buffers
//        (local L2 or peer HBM over NVLink) straight into registers
//     5. k more iterations in registers
//
// Two buffer parities make the strips double-buffered: writing version b never races with a
// neighbour still reading version b-1, and passing step 3 of exchange b proves every neighbour is
// done with version b-1. The final interior goes to the parity that holds no live strips.
// Split-phase variant (resident2_kernel.cuh, params.dual = 1): the exchange above is *exposed* — every
// tile needs its neighbours' block-b state exactly when it finishes block b itself. The split-phase
// kernel gives each CTA two vertically stacked virtual tiles and alternates between them; CTA rows
// alternate the order (even rows top-first, odd rows bottom-first), so virtual tile rows are phased
// F S S F F S S F ... While one tile computes, the other tile's registers are dead, so its halo is
// re-loaded mid-compute straight into them and lands in the background; every inter-CTA dependency
// then has at least a quarter period of slack. What stays zero-slack is the part of an F tile's halo
// that faces its sibling row (sibling: same CTA, program order; the two lateral corner pieces: flags).
// Deadlock-free by induction: F publishes of epoch b depend on epoch b-1 only, S publishes of epoch b
// on F publishes of epoch b. The WAR argument of the parity double buffer is unchanged (a tile's flag
// for epoch b+1 is set after both parts of its reload for block b were consumed).
//
// This is the north-star design of BASELINE.json: persistent, double-buffered, P2P loads of
// neighbour borders inside the stencil kernel, no NCCL/MPI on the path. It replaces the reference's
// per-iteration MPI_Isend/Irecv/Wait schedule (mpi/mpi_conv.c:164-248) and its launch-per-
// iteration CUDA loop (cuda/cuda_convolute.cu:66-87).
#include <cuda_runtime.h>

#include <cstring>

#include "resident2_kernel.cuh"
#include "resident3_kernel.cuh"
#include "resident_kernel.cuh"

namespace pconv {

// variant: 0 single-phase, 1 split-phase (resident2), 2 split-phase + protocol warp (resident3)
static ResidentFn pick_kernel(int kind, int C, int R, int variant = 0) {
    switch (kind) {
        case PCONV_KIND_SEP121:
            return variant == 2 ? pick_resident3_sep(C, R) : variant == 1 ? pick_resident2_sep(C, R) : pick_resident_sep(C, R);
        case PCONV_KIND_INT:
            return variant == 2 ? pick_resident3_int(C, R) : variant == 1 ? pick_resident2_int(C, R) : pick_resident_int(C, R);
        case PCONV_KIND_F32:
            return variant == 2 ? pick_resident3_f32(C, R) : variant == 1 ? pick_resident2_f32(C, R) : pick_resident_f32(C, R);
        default: return nullptr;
    }
}

}  // namespace pconv

using namespace pconv;

extern "C" int pconv_resident_launch(const pconv_resident_params_t* p, void* stream) {
    if (!p) return pconv_rt::fail("null params");
    const pconv_geometry_t& g = p->g;
    if (p->loops < 1 || p->k < 1 || p->TH < 1 || p->TWB < 16 || (p->TWB & 15) || (p->HB & 15) ||
        p->HB < p->k * g.C || p->TWB + 2 * p->HB > kSP ||
        p->TH + 2 * p->k > (p->dual == 2 ? kCompWarps : kWarps) * p->R || p->TH < p->k || p->TWB < p->HB)
        return pconv_rt::fail("bad resident geometry (loops/k/TH/TWB/HB/R)");
    const int64_t tile_rows = g.row_lo[2] - g.row_lo[1], tile_cb = g.cb_lo[2] - g.cb_lo[1];
    if ((int64_t)p->nty * p->TH < tile_rows || (int64_t)p->ntx * p->TWB < tile_cb)
        return pconv_rt::fail("CTA grid does not cover the tile");
    if (g.pitch <= 0 || g.pitch >= ((int64_t)1 << 30)) return pconv_rt::fail("tile pitch out of range");
    const bool dual = p->dual != 0;   // split-phase: nty counts virtual tile rows, two per CTA
    if (dual && (p->nty & 1)) return pconv_rt::fail("split-phase resident kernel needs an even number of tile rows");
    if ((int64_t)(p->nty + 2) * (p->ntx + 2) > 1024) return pconv_rt::fail("CTA flag grid exceeds its 4 KB pad");
    const int grid = (dual ? p->nty / 2 : p->nty) * p->ntx;
    if (grid < 1 || grid > pconv_rt::sm_count()) return pconv_rt::fail("resident grid exceeds the SM count");
    for (int s = 0; s < 3; s += 2) {
        const int64_t nr = g.row_lo[s + 1] - g.row_lo[s], nc = g.cb_lo[s + 1] - g.cb_lo[s];
        if ((nr > 0 && nr < p->k) || (nc > 0 && nc < (int64_t)p->k * g.C))
            return pconv_rt::fail("halo depth exceeds a neighbouring tile");
    }
    // per-CTA flags only cover the eight surrounding CTAs (see planner.cpp)
    if (g.row_lo[3] > g.row_lo[2] && tile_rows - (int64_t)(p->nty - 1) * p->TH < p->k)
        return pconv_rt::fail("last CTA row is shallower than the halo and a rank lies below");
    if (g.cb_lo[3] > g.cb_lo[2] && tile_cb - (int64_t)(p->ntx - 1) * p->TWB < (int64_t)p->k * g.C)
        return pconv_rt::fail("last CTA column is narrower than the halo and a rank lies to the right");
    ResidentFn fn = pick_kernel(p->filt.kind, g.C, p->R, (int)p->dual);
    if (!fn) return pconv_rt::fail("unsupported (filter kind, channels, band height) combination");

    ResidentArgs a;
    std::memset(&a, 0, sizeof(a));
    a.g = g;
    for (int d = 0; d < 9; ++d) {
        a.src[0].p[d] = p->src_a[d]; a.src[0].pitch[d] = g.pitch;
        a.src[1].p[d] = p->src_b[d]; a.src[1].pitch[d] = g.pitch;
        a.remote_cta_flags[d] = p->remote_cta_flags[d];
    }
    a.self[0] = p->self_a; a.self[1] = p->self_b;
    pconv_rt::make_filter_dev(p->filt, &a.f);
    a.loops = p->loops; a.k = p->k; a.TH = p->TH; a.TWB = p->TWB; a.HB = p->HB;
    a.nty = p->nty; a.ntx = p->ntx;
    a.cta_flags = p->cta_flags;
    a.epoch_base = p->epoch_base;
    a.sync = p->sync;
    if (a.sync.timeout_ns == 0) a.sync.timeout_ns = 10ull * 1000 * 1000 * 1000;

    void* args[] = {(void*)&a};
    cudaError_t e = cudaLaunchCooperativeKernel((const void*)fn, dim3(grid), dim3(kThreads), args, 0,
                                                (cudaStream_t)stream);
    pconv_rt::count_launch();
    if (e != cudaSuccess) return pconv_rt::fail_cuda(e, "resident_kernel launch");
    return 0;
}

extern "C" int pconv_resident2_band_heights(int kind, int* out, int cap) {
    int n = 0;
#define PUSH(RR) if (n < cap) out[n] = RR; ++n;
    if (kind == PCONV_KIND_F32) { PCONV_RES2_R_F32(PUSH) }
    else if (kind == PCONV_KIND_INT) { PCONV_RES2_R_INT(PUSH) }
    else { PCONV_RES2_R_SEP(PUSH) }
#undef PUSH
    return n;
}

extern "C" int pconv_resident3_band_heights(int kind, int* out, int cap) {
    int n = 0;
#define PUSH(RR) if (n < cap) out[n] = RR; ++n;
    if (kind == PCONV_KIND_F32) { PCONV_RES3_R_F32(PUSH) }
    else if (kind == PCONV_KIND_INT) { PCONV_RES3_R_INT(PUSH) }
    else { PCONV_RES3_R_SEP(PUSH) }
#undef PUSH
    return n;
}

extern "C" int pconv_resident_prepare(int kind, int C, int R, int variant) {
    ResidentFn fn = pick_kernel(kind, C, R, variant);
    if (!fn) return pconv_rt::fail("unsupported (filter kind, channels, band height) combination");
    cudaFuncAttributes at;
    cudaError_t e = cudaFuncGetAttributes(&at, (const void*)fn);
    if (e != cudaSuccess) return pconv_rt::fail_cuda(e, "cudaFuncGetAttributes");
    return 0;
}
