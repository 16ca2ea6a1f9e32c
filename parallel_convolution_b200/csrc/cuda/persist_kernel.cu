// Launcher of the persistent kernel (persist_kernel.cuh): validation of the plan, exchange-arena
// layout (common/xplan.h), dynamic shared memory, cooperative launch (every CTA must be co-resident:
// they wait on each other's flags).
#include <cuda_runtime.h>

#include <cstring>
#include <map>
#include <mutex>
#include <tuple>

#include "persist_kernel.cuh"

namespace pconv {

static PersistFn pick_kernel(int kind, int C, int R, int W) {
    if (W == 8) {
        switch (kind) {
            case PCONV_KIND_SEP121: return pick_persist_sep_w8(C, R);
            case PCONV_KIND_INT: return pick_persist_int_w8(C, R);
            case PCONV_KIND_F32: return pick_persist_f32_w8(C, R);
        }
    } else if (W == 16) {
        switch (kind) {
            case PCONV_KIND_SEP121: return pick_persist_sep_w16(C, R);
            case PCONV_KIND_INT: return pick_persist_int_w16(C, R);
            case PCONV_KIND_F32: return pick_persist_f32_w16(C, R);
        }
    }
    return nullptr;
}

static size_t smem_bytes(int W, const XGeom& x) {
    const size_t head = W == 8 ? offsetof(PersistSmem<8>, inbox) : offsetof(PersistSmem<16>, inbox);
    return head + (size_t)x.slot_chunks * 16 + (size_t)2 * x.k * 32 * 16;   // + outbox
}

// Geometry checks shared by the launcher and pconv_persist_arena_bytes; fills *x.
static const char* make_xgeom(const pconv_persist_params_t& p, XGeom* x) {
    const pconv_geometry_t& g = p.g;
    const int64_t tile_rows = g.row_lo[2] - g.row_lo[1], tile_cb = g.cb_lo[2] - g.cb_lo[1];
    if (p.W != 8 && p.W != 16) return "warps per CTA must be 8 or 16";
    if (p.loops < 1 || p.k < 1 || p.TH < 1 || p.R < 1 || p.TWB < 16 || (p.TWB & 15) || (p.HB & 15) || p.HB < 16 ||
        p.HB < p.k * g.C || p.TWB + 2 * p.HB != kSP || p.TH + 2 * p.k > p.W * p.R || p.TH < p.k || p.TWB < p.HB)
        return "bad persistent-kernel geometry (loops/k/TH/TWB/HB/R/W)";
    if (tile_rows < 1 || tile_cb < 1 || tile_rows >= (1 << 30) || tile_cb >= (1 << 30)) return "tile extent out of range";
    if ((int64_t)p.nty * p.TH < tile_rows || (int64_t)(p.nty - 1) * p.TH >= tile_rows ||
        (int64_t)p.ntx * p.TWB < tile_cb || (int64_t)(p.ntx - 1) * p.TWB >= tile_cb)
        return "CTA grid does not match the tile";
    if ((int64_t)p.nty * p.ntx > kXMaxCtas) return "CTA grid exceeds the flag area";
    if (g.pitch <= 0 || g.pitch >= ((int64_t)1 << 30)) return "tile pitch out of range";
    for (int s = 0; s < 3; s += 2) {
        const int64_t nr = g.row_lo[s + 1] - g.row_lo[s], nc = g.cb_lo[s + 1] - g.cb_lo[s];
        if ((nr > 0 && nr < p.k) || (nc > 0 && nc < (int64_t)p.k * g.C)) return "halo depth exceeds a neighbouring tile";
    }
    // A CTA exchanges with its eight surrounding CTAs only: with a rank below / to the right, the last
    // CTA row / column must itself be one halo deep (and whole chunks wide), or the halo of the row /
    // column before it would reach past it.
    const int64_t last_rows = tile_rows - (int64_t)(p.nty - 1) * p.TH, last_cb = tile_cb - (int64_t)(p.ntx - 1) * p.TWB;
    if (g.row_lo[3] > g.row_lo[2] && last_rows < p.k) return "last CTA row is shallower than the halo and a rank lies below";
    if (g.cb_lo[3] > g.cb_lo[2] && (last_cb < p.HB || (tile_cb & 15)))
        return "last CTA column is narrower than the halo (or not whole chunks) and a rank lies to the right";
    if (g.cb_lo[1] > g.cb_lo[0] && (tile_cb & 15)) return "tile width must be whole 16-byte chunks when a rank lies to the left";
    std::memset(x, 0, sizeof(*x));
    x->nty = p.nty; x->ntx = p.ntx; x->TH = p.TH; x->TWB = p.TWB; x->HB = p.HB; x->k = p.k; x->R = p.R; x->W = p.W;
    x->tile_rows = (int32_t)tile_rows; x->tile_cb = (int32_t)tile_cb;
    xp_finish(x);
    return nullptr;
}

}  // namespace pconv

using namespace pconv;

extern "C" int64_t pconv_persist_arena_bytes(const pconv_persist_params_t* p) {
    XGeom x;
    if (!p) return pconv_rt::fail("null params");
    if (const char* why = make_xgeom(*p, &x)) return pconv_rt::fail(why);
    return xp_arena_bytes(x);
}

extern "C" int pconv_persist_band_heights(int kind, int* out, int cap) {
    int n = 0;
#define PUSH(RR) { if (n < cap) out[n] = RR; ++n; }
    if (kind == PCONV_KIND_F32) { PCONV_PER_R_F32(PUSH) }
    else if (kind == PCONV_KIND_INT) { PCONV_PER_R_INT(PUSH) }
    else { PCONV_PER_R_SEP(PUSH) }
#undef PUSH
    return n;
}

// CTAs of this (kind, C, R, W) instance that fit on the device at once with the plan's shared memory;
// also forces the lazy module load. <0 on error.
extern "C" int pconv_persist_capacity(const pconv_persist_params_t* p) {
    XGeom x;
    if (!p) return pconv_rt::fail("null params");
    if (const char* why = make_xgeom(*p, &x)) return pconv_rt::fail(why);
    PersistFn fn = pick_kernel(p->filt.kind, p->g.C, p->R, p->W);
    if (!fn) return pconv_rt::fail("unsupported (filter kind, channels, band height, warps) combination");
    const size_t smem = smem_bytes(p->W, x);
    if ((int)smem > pconv_rt::max_optin_smem()) return pconv_rt::fail("inbox does not fit in shared memory");
    // per (device, kernel instance, shared-memory size): the attribute and the occupancy query cost tens of
    // microseconds of host time — too much to repeat on every launch of a 200 us step
    static std::mutex mu;
    static std::map<std::tuple<int, const void*, size_t>, int> cache;     // -> resident CTAs
    static std::map<std::pair<int, const void*>, size_t> limit;           // largest dynamic smem allowed so far
    int dev = 0;
    cudaGetDevice(&dev);
    const auto key = std::make_tuple(dev, (const void*)fn, smem);
    std::lock_guard<std::mutex> lk(mu);
    auto it = cache.find(key);
    if (it != cache.end()) return it->second;
    size_t& lim = limit[std::make_pair(dev, (const void*)fn)];
    if (smem > lim) {      // only ever RAISE the limit: a cached smaller plan must not lower it under a larger one
        cudaError_t e = cudaFuncSetAttribute((const void*)fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return pconv_rt::fail_cuda(e, "cudaFuncSetAttribute(persist_kernel)");
        lim = smem;
    }
    int per_sm = 0;
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, (const void*)fn, p->W * 32, smem);
    if (e != cudaSuccess) return pconv_rt::fail_cuda(e, "cudaOccupancyMaxActiveBlocksPerMultiprocessor");
    const int cap = per_sm * pconv_rt::sm_count();
    cache[key] = cap;
    return cap;
}

extern "C" int pconv_persist_launch(const pconv_persist_params_t* p, void* stream) {
    if (!p) return pconv_rt::fail("null params");
    PersistArgs a;
    std::memset(&a, 0, sizeof(a));
    if (const char* why = make_xgeom(*p, &a.x)) return pconv_rt::fail(why);
    if (!p->arena[4]) return pconv_rt::fail("no exchange arena");
    if (p->arena_bytes < xp_arena_bytes(a.x)) return pconv_rt::fail("exchange arena too small for this plan");
    PersistFn fn = pick_kernel(p->filt.kind, p->g.C, p->R, p->W);
    if (!fn) return pconv_rt::fail("unsupported (filter kind, channels, band height, warps) combination");
    const int cap = pconv_persist_capacity(p);
    if (cap < 0) return cap;
    const int grid = p->nty * p->ntx;
    if (grid > cap) return pconv_rt::fail("persistent grid exceeds what the device can keep resident");

    a.g = p->g;
    for (int d = 0; d < 9; ++d) {
        a.src.p[d] = p->src[d];
        a.src.pitch[d] = p->g.pitch;
        a.arena[d] = p->arena[d];
    }
    a.dst = p->dst;
    pconv_rt::make_filter_dev(p->filt, &a.f);
    a.loops = p->loops;
    a.epoch_base = p->epoch_base;
    a.peer_scope = p->peer_scope;
    a.debug_skip = p->debug_skip;
    a.sync = p->sync;
    if (a.sync.timeout_ns == 0) a.sync.timeout_ns = 10ull * 1000 * 1000 * 1000;

    void* args[] = {(void*)&a};
    cudaError_t e = cudaLaunchCooperativeKernel((const void*)fn, dim3(grid), dim3(p->W * 32), args,
                                                smem_bytes(p->W, a.x), (cudaStream_t)stream);
    pconv_rt::count_launch();
    if (e != cudaSuccess) return pconv_rt::fail_cuda(e, "persist_kernel launch");
    return 0;
}
