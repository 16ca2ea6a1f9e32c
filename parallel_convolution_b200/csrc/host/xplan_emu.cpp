// CPU emulation of the persistent kernel's exchange protocol (csrc/cuda/persist_kernel.cuh).
//
// Every (rank, CTA, warp, lane) of a launch is replayed sequentially with the SAME index functions
// the device code uses (common/xplan.h: xp_thread / xp_push / xp_fetch / xp_peer / slot layout):
// regions are byte arrays instead of registers, the inbox slots are host arrays laid out exactly like
// the device arena, the bulk copies are memcpy of the same (offset, size) pairs, and one stencil
// iteration is the oracle's arithmetic on the region (zero beyond the region edge — that ring is
// stale on the device as well and must never reach the interior within k iterations).
// Unwritten inbox bytes are poisoned, so a block that is pushed to the wrong place, not pushed at
// all, or fetched with the wrong size shows up as a mismatch against the whole-image oracle.
// This is how the mailbox geometry (ragged last rows / columns, 8 neighbours, cross-rank wrap of the
// CTA coordinates) is tested on a machine without a GPU; the reference has no counterpart (its halo
// exchange was only ever checked by eye).
#include <cstdint>
#include <cstring>
#include <vector>

#include "../common/xplan.h"
#include "pconv_host.h"

namespace {

using namespace pconv;

inline uint8_t saturate_trunc(float v) {
    if (!(v > 0.0f)) return 0;
    if (v >= 255.0f) return 255;
    return (uint8_t)(int)v;
}

struct Cta {
    std::vector<uint8_t> reg;   // RH x 512
};

// one iteration on a region: byte stream with channel stride C, zero beyond the region, then the
// image mask (bytes outside the image stay zero: zero padding)
void region_pass(std::vector<uint8_t>& reg, std::vector<uint8_t>& tmp, int RH, int C, const float* h,
                 int64_t gr0, int64_t gb0, int64_t img_rows, int64_t img_cb) {
    for (int r = 0; r < RH; ++r) {
        for (int xb = 0; xb < 512; ++xb) {
            const int64_t gr = gr0 + r, gb = gb0 + xb;
            if (gr < 0 || gr >= img_rows || gb < 0 || gb >= img_cb) { tmp[(size_t)r * 512 + xb] = 0; continue; }
            volatile float val = 0.0f;
            for (int ky = 0; ky < 3; ++ky)
                for (int kx = 0; kx < 3; ++kx) {
                    const int rr = r + ky - 1, xx = xb + (kx - 1) * C;
                    float p = 0.0f;
                    if (rr >= 0 && rr < RH && xx >= 0 && xx < 512) p = (float)reg[(size_t)rr * 512 + xx];
                    volatile float prod = p * h[3 * ky + kx];
                    val = val + prod;
                }
            tmp[(size_t)r * 512 + xb] = saturate_trunc(val);
        }
    }
    reg.swap(tmp);
}

}  // namespace

// img/out: H x (Wpx*C) dense. The image is split into gy x gx equal tiles (must divide exactly), each
// covered by CTAs of W warps x R rows with halo depth k and TH interior rows per CTA (0: as many as
// fit, balanced). Returns 0, or a negative code for a geometry the device launcher would reject:
// -2 k/R/W do not fit, -3 tiles not equal, -4 last row/column rules, -5 chunk alignment.
extern "C" int pconv_xplan_emulate(const uint8_t* img, uint8_t* out, int64_t H, int64_t Wpx, int C,
                                   const pconv_filter_t* f, int loops, int gy, int gx, int k, int R, int W,
                                   int TH, uint32_t epoch_base) {
    if (!img || !out || !f || H < 1 || Wpx < 1 || C < 1 || loops < 1 || gy < 1 || gx < 1 || k < 1) return -1;
    const int64_t img_cb = Wpx * C;
    if (H % gy || Wpx % gx) return -3;
    const int64_t tile_rows = H / gy, tile_cb = img_cb / gx;
    XGeom x;
    std::memset(&x, 0, sizeof(x));
    x.k = k; x.R = R; x.W = W;
    x.HB = (k * C + 15) / 16 * 16;
    x.TWB = 512 - 2 * x.HB;
    const int RH = W * R;
    if (x.TWB < x.HB || RH - 2 * k < k) return -2;
    int nty = (int)((tile_rows + (RH - 2 * k) - 1) / (RH - 2 * k));
    if (TH <= 0) TH = (int)((tile_rows + nty - 1) / nty);
    if (TH < k) TH = k;
    if (TH + 2 * k > RH) return -2;
    nty = (int)((tile_rows + TH - 1) / TH);
    x.TH = TH;
    x.nty = nty;
    x.ntx = (int)((tile_cb + x.TWB - 1) / x.TWB);
    x.tile_rows = (int32_t)tile_rows; x.tile_cb = (int32_t)tile_cb;
    xp_finish(&x);
    const int64_t last_rows = tile_rows - (int64_t)(x.nty - 1) * x.TH, last_cb = tile_cb - (int64_t)(x.ntx - 1) * x.TWB;
    if (gy > 1 && (last_rows < k || tile_rows < k)) return -4;
    if (gx > 1 && (last_cb < x.HB || tile_cb < (int64_t)k * C)) return -4;
    if (gx > 1 && (tile_cb & 15)) return -5;
    if ((int64_t)x.nty * x.ntx > kXMaxCtas) return -2;

    const int nranks = gy * gx, ncta = x.nty * x.ntx;
    std::vector<std::vector<uint8_t>> arena(nranks, std::vector<uint8_t>((size_t)xp_arena_bytes(x), 0));
    std::vector<std::vector<Cta>> ctas(nranks, std::vector<Cta>(ncta));
    auto rank_in_dir = [&](int rank, int rd) -> int {
        const int ry = rank / gx + rd / 3 - 1, rx = rank % gx + rd % 3 - 1;
        return (ry < 0 || ry >= gy || rx < 0 || rx >= gx) ? -1 : ry * gx + rx;
    };
    auto origin = [&](int rank, int c, int64_t* gr0, int64_t* gb0) {
        const int ti = c / x.ntx, tj = c % x.ntx;
        *gr0 = (rank / gx) * tile_rows + (int64_t)ti * x.TH - k;
        *gb0 = (rank % gx) * tile_cb + (int64_t)tj * x.TWB - x.HB;
    };
    // initial region load
    for (int r = 0; r < nranks; ++r)
        for (int c = 0; c < ncta; ++c) {
            auto& reg = ctas[r][c].reg;
            reg.assign((size_t)RH * 512, 0);
            int64_t gr0, gb0;
            origin(r, c, &gr0, &gb0);
            for (int rr = 0; rr < RH; ++rr)
                for (int xb = 0; xb < 512; ++xb) {
                    const int64_t gr = gr0 + rr, gb = gb0 + xb;
                    if (gr >= 0 && gr < H && gb >= 0 && gb < img_cb) reg[(size_t)rr * 512 + xb] = img[gr * img_cb + gb];
                }
        }
    std::vector<uint8_t> tmp((size_t)RH * 512), inbox((size_t)x.slot_chunks * 16);
    const int nblocks = (loops + k - 1) / k;
    for (int b = 0; b < nblocks; ++b) {
        if (b > 0) {
            const uint32_t epoch = epoch_base + (uint32_t)b;
            const int par = (int)(epoch & 1u);
            // poison this parity of every slot, then PUSH from every thread of every CTA
            for (int r = 0; r < nranks; ++r)
                std::memset(arena[r].data() + xp_slot_byte(x, 0, par), 0xCD, (size_t)ncta * x.slot_chunks * 16);
            for (int r = 0; r < nranks; ++r)
                for (int c = 0; c < ncta; ++c) {
                    const int ti = c / x.ntx, tj = c % x.ntx, vr = xp_vr(x, ti), vc = xp_vc(x, tj);
                    // vertical step 1 (device: border warps dump their rows into the shared-memory outbox)
                    std::vector<uint8_t> outbox((size_t)2 * k * 512, 0xAB);
                    for (int w = 0; w < W; ++w)
                        for (int lane = 0; lane < 32; ++lane) {
                            const XThread t = xp_thread(x, vr, vc, w * R, lane, R);
                            for (int i = 0; i < R; ++i) {
                                const uint8_t* src = &ctas[r][c].reg[(size_t)(w * R + i) * 512 + 16 * lane];
                                if ((t.m_up >> i) & 1u) std::memcpy(&outbox[((size_t)(w * R + i - k) * 32 + lane) * 16], src, 16);
                                if ((t.m_dn >> i) & 1u) std::memcpy(&outbox[((size_t)(k + w * R + i - vr) * 32 + lane) * 16], src, 16);
                            }
                        }
                    for (int D = 0; D < 9; ++D) {
                        if (D == 4) continue;
                        const XPeer p = xp_peer(x, ti, tj, D);
                        const int dst_rank = p.rank_dir == 4 ? r : rank_in_dir(r, p.rank_dir);
                        if (dst_rank < 0) continue;
                        uint8_t* slot = arena[dst_rank].data() + xp_slot_byte(x, p.cta, par);
                        if (D == 3 || D == 5) {      // horizontal: straight from the "registers", per thread
                            for (int w = 0; w < W; ++w)
                                for (int lane = 0; lane < 32; ++lane) {
                                    const XThread t = xp_thread(x, vr, vc, w * R, lane, R);
                                    uint32_t m;
                                    int idx0, stride;
                                    if (!xp_push(x, t, D, vr, vc, w * R, lane, &m, &idx0, &stride)) continue;
                                    for (int i = 0; i < R; ++i)
                                        if ((m >> i) & 1u) {
                                            const int idx = idx0 + i * stride;
                                            if (idx < 0 || idx >= x.slot_chunks) return -10;   // out of the slot
                                            std::memcpy(slot + (size_t)idx * 16, &ctas[r][c].reg[(size_t)(w * R + i) * 512 + 16 * lane], 16);
                                        }
                                }
                        } else {                     // vertical step 2: (direction, row) pairs from the outbox
                            const XBlock sb = xp_send(x, D, vr, vc);
                            for (int rr = 0; rr < k; ++rr)
                                for (int lane = 0; lane < sb.width; ++lane) {
                                    const int idx = x.off[8 - D] + rr * sb.width + lane;
                                    if (idx < 0 || idx >= x.slot_chunks || sb.chunk0 + lane >= 32) return -10;
                                    std::memcpy(slot + (size_t)idx * 16,
                                                &outbox[((size_t)((D > 4 ? k : 0) + rr) * 32 + sb.chunk0 + lane) * 16], 16);
                                }
                        }
                    }
                }
            // FETCH: bulk copies of the existing sources' blocks, then per-thread halo refresh
            for (int r = 0; r < nranks; ++r)
                for (int c = 0; c < ncta; ++c) {
                    const int ti = c / x.ntx, tj = c % x.ntx, vr = xp_vr(x, ti), vc = xp_vc(x, tj);
                    const uint8_t* slot = arena[r].data() + xp_slot_byte(x, c, par);
                    std::memset(inbox.data(), 0xEE, inbox.size());
                    uint32_t nbr_mask = 0;
                    for (int s = 0; s < 9; ++s) {
                        if (s == 4) continue;
                        const XPeer p = xp_peer(x, ti, tj, s);
                        if (p.rank_dir != 4 && rank_in_dir(r, p.rank_dir) < 0) continue;
                        nbr_mask |= 1u << s;
                        const XBlock blk = xp_block(x, s, vr, vc);
                        const size_t bytes = (size_t)blk.nrows * blk.width * 16;
                        if ((size_t)x.off[s] * 16 + bytes > inbox.size()) return -11;
                        std::memcpy(inbox.data() + (size_t)x.off[s] * 16, slot + (size_t)x.off[s] * 16, bytes);
                    }
                    for (int w = 0; w < W; ++w)
                        for (int lane = 0; lane < 32; ++lane) {
                            const XThread t = xp_thread(x, vr, vc, w * R, lane, R);
                            for (int grp = 0; grp < 3; ++grp) {
                                uint32_t m;
                                int src, idx0, stride;
                                if (!xp_fetch(x, t, grp, vr, vc, w * R, lane, &src, &m, &idx0, &stride)) continue;
                                if (!((nbr_mask >> src) & 1u)) continue;
                                for (int i = 0; i < R; ++i)
                                    if ((m >> i) & 1u) {
                                        const int idx = idx0 + i * stride;
                                        if (idx < 0 || idx >= x.slot_chunks) return -12;
                                        std::memcpy(&ctas[r][c].reg[(size_t)(w * R + i) * 512 + 16 * lane], inbox.data() + (size_t)idx * 16, 16);
                                    }
                            }
                        }
                }
        }
        const int kb = loops - b * k < k ? loops - b * k : k;
        for (int r = 0; r < nranks; ++r)
            for (int c = 0; c < ncta; ++c) {
                int64_t gr0, gb0;
                origin(r, c, &gr0, &gb0);
                for (int it = 0; it < kb; ++it) region_pass(ctas[r][c].reg, tmp, RH, C, f->h, gr0, gb0, H, img_cb);
            }
    }
    // final interior store
    for (int r = 0; r < nranks; ++r)
        for (int c = 0; c < ncta; ++c) {
            const int ti = c / x.ntx, tj = c % x.ntx, vr = xp_vr(x, ti);
            int64_t gr0, gb0;
            origin(r, c, &gr0, &gb0);
            const int64_t cb_lo = (r % gx) * tile_cb + (int64_t)tj * x.TWB;
            int64_t cb_hi = cb_lo + x.TWB;
            if (cb_hi > (r % gx + 1) * tile_cb) cb_hi = (r % gx + 1) * tile_cb;
            for (int rr = k; rr < k + vr; ++rr)
                for (int64_t gb = cb_lo; gb < cb_hi; ++gb)
                    out[(gr0 + rr) * img_cb + gb] = ctas[r][c].reg[(size_t)rr * 512 + (gb - gb0)];
        }
    return 0;
}
