// Exchange plan of the persistent kernel — pure index arithmetic shared by the device code
// (csrc/cuda/persist_kernel.cuh), the launcher, the planner and a CPU emulation of the whole
// protocol (csrc/host/xplan_emu.cpp, run by the CPU test-suite: the mailbox geometry is validated
// against the oracle without a GPU).
//
// Geometry. A rank's tile (tile_rows x tile_cb bytes) is covered by nty x ntx CTAs. CTA (ti,tj)
// holds a *region* of W*R rows x 32 chunks (a chunk = 16 bytes = one lane's strip of a row):
//
//      rows  [0,k)            top halo          chunks [0,nh)            left halo (nh = HB/16)
//            [k,k+vr)         interior                 [nh,nh+vc)        interior
//            [k+vr,2k+vr)     bottom halo              [nh+vc,2nh+vc)    right halo
//
// vr/vc are TH rows / TWB/16 chunks except in the last CTA row / column of the tile. Every k
// iterations a CTA *pushes* the k-deep border of its interior into the inboxes of its (up to) eight
// neighbouring CTAs — CTAs of the same GPU, or CTAs of the neighbouring rank's GPU through a
// peer-mapped pointer — and announces it with one release-store per neighbour. An inbox is one
// contiguous slot of global memory per (CTA, parity) that lives on the CONSUMER's GPU, made of one
// dense block per source direction, so that the consumer fetches each block with a single bulk copy
// (cp.async.bulk -> shared memory, completing on an mbarrier) as soon as that neighbour's flag is
// seen. All of a consumer's polling and halo reads are local; only stores cross NVLink.
//
// The reference moves the same data with MPI_Isend/Irecv + MPI_Type_vector, four neighbours only
// (mpi/mpi_conv.c:80-88,166-200; the diagonal cells are never exchanged).
//
// Directions are row-major (dy,dx) in {-1,0,1}^2: 0=NW 1=N 2=NE 3=W 4=self 5=E 6=SW 7=S 8=SE.
#ifndef PCONV_XPLAN_H
#define PCONV_XPLAN_H

#include <stdint.h>

#if defined(__CUDACC__)
#define XP_HD __host__ __device__ __forceinline__
#else
#define XP_HD inline
#endif

namespace pconv {

constexpr int kXFlagStride = 16;     // u32 words per CTA in the flag array (one 64-byte line)
constexpr int kXMaxCtas = 1024;      // CTAs per rank the flag area is sized for
constexpr int64_t kXFlagBytes = (int64_t)kXMaxCtas * kXFlagStride * 4;

struct XGeom {
    int32_t nty, ntx;        // CTA grid over the tile (identical on every rank)
    int32_t TH, TWB, HB, k;  // interior rows / bytes per CTA, column halo bytes, halo depth
    int32_t R, W;            // band height (rows per warp), warps per CTA
    int32_t tile_rows, tile_cb;
    int32_t nh, NC;          // HB/16, TWB/16
    int32_t slot_chunks;     // inbox slot size (one CTA, one parity) in 16-byte chunks
    int32_t off[9];          // chunk offset of the block received FROM direction s inside a slot
};

XP_HD int xp_vr(const XGeom& x, int ti) {
    const int left = x.tile_rows - ti * x.TH;
    return left < x.TH ? left : x.TH;
}
XP_HD int xp_vc(const XGeom& x, int tj) {
    int left = x.tile_cb - tj * x.TWB;
    if (left > x.TWB) left = x.TWB;
    return (left + 15) >> 4;
}

// Fill the derived fields (nh, NC, off[], slot_chunks). Block capacities use the full-size CTA so
// that the offsets are the same for every CTA of every rank.
inline void xp_finish(XGeom* x) {
    x->nh = x->HB / 16;
    x->NC = x->TWB / 16;
    const int cap_v = x->k * x->NC, cap_h = x->TH * x->nh, cap_c = x->k * x->nh;
    int o = 0;
    const int order[8] = {1, 7, 3, 5, 0, 2, 6, 8};
    const int cap[8] = {cap_v, cap_v, cap_h, cap_h, cap_c, cap_c, cap_c, cap_c};
    for (int i = 0; i < 8; ++i) { x->off[order[i]] = o; o += cap[i]; }
    x->off[4] = 0;
    x->slot_chunks = (o + 7) & ~7;   // 128-byte multiple
}

// Block received from direction s by a consumer with (vr, vc): where it sits in the region.
struct XBlock {
    int row0, nrows;     // region rows [row0, row0+nrows)
    int chunk0, width;   // region chunks [chunk0, chunk0+width); the block is row-major, `width` wide
};
XP_HD XBlock xp_block(const XGeom& x, int s, int vr, int vc) {
    const int sy = s / 3 - 1, sx = s % 3 - 1;
    XBlock b;
    b.row0 = sy < 0 ? 0 : (sy == 0 ? x.k : x.k + vr);
    b.nrows = sy == 0 ? vr : x.k;
    b.chunk0 = sx < 0 ? 0 : (sx == 0 ? x.nh : x.nh + vc);
    b.width = sx == 0 ? vc : x.nh;
    return b;
}

// What a producer with (vr, vc) sends towards direction D: the part of its own region that becomes
// the consumer's block "from direction 8-D" (same nrows x width, row-major).
XP_HD XBlock xp_send(const XGeom& x, int D, int vr, int vc) {
    const int dy = D / 3 - 1, dx = D % 3 - 1;
    XBlock b;
    b.row0 = dy < 0 ? x.k : (dy == 0 ? x.k : vr);   // dy>0: the last k interior rows [k+vr-k, k+vr)
    b.nrows = dy == 0 ? vr : x.k;
    b.chunk0 = dx < 0 ? x.nh : (dx == 0 ? x.nh : vc);   // dx>0: the last nh interior chunks
    b.width = dx == 0 ? vc : x.nh;
    return b;
}

// The CTA reached from (ti,tj) in direction D: rank direction it lives in (4 = this rank) and its
// CTA index there.
struct XPeer {
    int rank_dir;   // 0..8
    int cta;        // ti'*ntx + tj'
};
XP_HD XPeer xp_peer(const XGeom& x, int ti, int tj, int D) {
    int ni = ti + D / 3 - 1, nj = tj + D % 3 - 1;
    int ry = 1, rx = 1;
    if (ni < 0) { ry = 0; ni += x.nty; } else if (ni >= x.nty) { ry = 2; ni -= x.nty; }
    if (nj < 0) { rx = 0; nj += x.ntx; } else if (nj >= x.ntx) { rx = 2; nj -= x.ntx; }
    XPeer p;
    p.rank_dir = ry * 3 + rx;
    p.cta = ni * x.ntx + nj;
    return p;
}

// ---- per-thread view ---------------------------------------------------------------------------
// Thread (warp, lane) holds band rows [rr0, rr0+R) (rr0 = warp*R) of chunk `lane`.
struct XThread {
    uint32_t m_top, m_int, m_bot;   // band-row masks (bit i = band row i): top halo / interior / bottom halo
    uint32_t m_up, m_dn;            // rows pushed upwards (N, NW, NE) / downwards (S, SW, SE)
    int32_t hcls;                   // lane class: 0 left halo, 1 interior, 2 right halo, 3 unused
};
XP_HD XThread xp_thread(const XGeom& x, int vr, int vc, int rr0, int lane, int R) {
    XThread t;
    t.m_top = t.m_int = t.m_bot = t.m_up = t.m_dn = 0u;
    for (int i = 0; i < R; ++i) {
        const int rr = rr0 + i;
        if (rr < x.k) t.m_top |= 1u << i;
        if (rr >= x.k && rr < x.k + vr) t.m_int |= 1u << i;
        if (rr >= x.k + vr && rr < 2 * x.k + vr) t.m_bot |= 1u << i;
        if (rr >= x.k && rr < 2 * x.k) t.m_up |= 1u << i;
        if (rr >= vr && rr < vr + x.k) t.m_dn |= 1u << i;
    }
    t.hcls = lane < x.nh ? 0 : (lane < x.nh + vc ? 1 : (lane < 2 * x.nh + vc ? 2 : 3));
    return t;
}

// Push towards direction D. Returns false when this thread sends nothing that way; otherwise the band
// rows to send (*mask), the chunk index of band row 0 inside the consumer's slot (*idx0, may be
// "before" the block for rows outside the mask) and the row stride in chunks.
XP_HD bool xp_push(const XGeom& x, const XThread& t, int D, int vr, int vc, int rr0, int lane,
                   uint32_t* mask, int* idx0, int* stride) {
    const int dy = D / 3 - 1, dx = D % 3 - 1;
    *mask = dy < 0 ? t.m_up : (dy > 0 ? t.m_dn : t.m_int);
    // my interior chunks are [nh, nh+vc): the first nh go west, the last nh east, all of them north/south
    const bool ok = t.hcls == 1 && (dx < 0 ? lane < 2 * x.nh : (dx > 0 ? lane >= vc : true));
    const int row0 = dy > 0 ? vr : x.k, chunk0 = dx > 0 ? vc : x.nh;
    *stride = dx == 0 ? vc : x.nh;
    *idx0 = x.off[8 - D] + (rr0 - row0) * *stride + (lane - chunk0);
    return ok && *mask != 0u;
}

// Fetch group grp (0: top halo rows, 1: interior rows = side halos, 2: bottom halo rows). Returns false
// when the thread has nothing to refresh there; otherwise the source direction (*s: skip if that
// neighbour does not exist), the band rows, and the chunk index of band row 0 inside my own slot.
XP_HD bool xp_fetch(const XGeom& x, const XThread& t, int grp, int vr, int vc, int rr0, int lane,
                    int* s, uint32_t* mask, int* idx0, int* stride) {
    if (t.hcls == 3 || (grp == 1 && t.hcls == 1)) return false;
    *s = 3 * grp + t.hcls;
    *mask = grp == 0 ? t.m_top : (grp == 1 ? t.m_int : t.m_bot);
    const int row0 = grp == 0 ? 0 : (grp == 1 ? x.k : x.k + vr);
    const int chunk0 = t.hcls == 0 ? 0 : (t.hcls == 1 ? x.nh : x.nh + vc);
    *stride = t.hcls == 1 ? vc : x.nh;
    *idx0 = x.off[*s] + (rr0 - row0) * *stride + (lane - chunk0);
    return *mask != 0u;
}

// Arena layout (one per rank, identical everywhere): flags | inbox parity 0 | inbox parity 1.
XP_HD int64_t xp_flag_word(int cta, int s) { return (int64_t)cta * kXFlagStride + s; }
XP_HD int64_t xp_slot_byte(const XGeom& x, int cta, int parity) {
    const int64_t ncta = (int64_t)x.nty * x.ntx;
    return kXFlagBytes + ((int64_t)parity * ncta + cta) * (int64_t)x.slot_chunks * 16;
}
inline int64_t xp_arena_bytes(const XGeom& x) {
    return kXFlagBytes + 2 * (int64_t)x.nty * x.ntx * (int64_t)x.slot_chunks * 16;
}

}  // namespace pconv
#endif
