"""Property-based tests (hypothesis) of the pure logic: decomposition, planner, filter
classification, and the exactness claims the integer paths rest on."""
import numpy as np
from hypothesis import given, settings, strategies as st

from parallel_convolution_b200.models.filters import Filter
from parallel_convolution_b200.ops import host
from parallel_convolution_b200.parallel.decomposition import TileGrid
from parallel_convolution_b200.parallel.plan import plan_persist, plan_stream


@settings(max_examples=60, deadline=None)
@given(h=st.integers(1, 400), w=st.integers(1, 400), n=st.integers(1, 16), c=st.sampled_from([1, 3, 4]))
def test_tiles_partition_the_image(h, w, n, c):
    try:
        g = TileGrid.create(h, w, c, n)
    except ValueError:
        return                      # more workers than pixels along every factorisation
    seen = np.zeros((h, w), dtype=np.int32)
    for t in g.tiles():
        assert t.rows >= 1 and t.cols >= 1
        seen[t.row0:t.row0 + t.rows, t.col0:t.col0 + t.cols] += 1
    assert (seen == 1).all()
    rows = {t.rows for t in g.tiles()}
    assert max(rows) - min(rows) <= 1           # balanced remainder split
    if h % g.grid_rows == 0 and w % g.grid_cols == 0:
        assert g.exact


@settings(max_examples=60, deadline=None)
@given(h=st.integers(2, 300), w=st.integers(2, 300), gr=st.integers(1, 4), gc=st.integers(1, 4))
def test_neighbor_table_is_symmetric(h, w, gr, gc):
    if gr > h or gc > w:
        return
    g = TileGrid.create(h, w, 1, gr * gc, (gr, gc))
    for r in range(g.workers):
        for d, nb in enumerate(g.neighbors(r)):
            if nb >= 0:
                assert g.neighbors(nb)[8 - d] == r      # I am my neighbour's opposite neighbour


@settings(max_examples=80, deadline=None)
@given(rows=st.integers(1, 40000), cb=st.integers(1, 120000), c=st.sampled_from([1, 3, 4]),
       kind=st.sampled_from([0, 1, 2]), loops=st.integers(1, 600))
def test_stream_plans_are_always_launchable(rows, cb, c, kind, loops):
    p = plan_stream(rows, cb * c, c, kind, loops)
    assert 1 <= p.k <= loops
    assert p.TH >= 1 and p.TH + 2 * p.k <= 16 * p.R
    assert p.HB % 16 == 0 and p.HB >= p.k * c and p.TWB + 2 * p.HB == 512
    assert p.ntx * p.TWB >= cb * c and p.nty * p.TH >= rows
    assert sum(p.launches(loops)) == loops and all(1 <= k <= p.k for k in p.launches(loops))
    for warps in (8, 16):
        r = plan_persist(rows, cb * c, c, kind, loops, warps=warps)
        if r is not None:
            assert r.W == warps and r.ntx * r.nty <= 148 * (16 // warps) and r.TH >= r.k and r.TWB >= r.HB


@settings(max_examples=40, deadline=None)
@given(w=st.lists(st.integers(-14, 14), min_size=9, max_size=9), s=st.integers(0, 7),
       seed=st.integers(0, 2**31 - 1))
def test_integer_filters_are_exact_in_fp32(w, s, seed):
    """The classifier's claim: integer taps over 2^s (sum|w| <= 128) => the fp32 oracle equals
    clamp(floor(S / 2^s), 0, 255) for every input."""
    f = Filter.from_int(w, float(1 << s))
    if f.kind_name == "f32-exact":
        assert sum(abs(v) for v in w) > 128
        return
    rng = np.random.default_rng(seed)
    img = rng.integers(0, 256, (12, 14), dtype=np.uint8)
    p = np.zeros((14, 16), dtype=np.int64)
    p[1:-1, 1:-1] = img
    wm = np.array(w).reshape(3, 3)
    S = sum(wm[k, l] * p[k:k + 12, l:l + 14] for k in range(3) for l in range(3))
    want = np.clip(S >> s, 0, 255).astype(np.uint8)
    assert np.array_equal(host.oracle(img, f.spec, 1), want)
    assert np.array_equal(host.cpu_run_image(img, f.spec, 1), want)


@settings(max_examples=150, deadline=None)
@given(rows=st.integers(40, 4000), cb=st.integers(64, 12000), c=st.sampled_from([1, 3, 4]),
       kind=st.sampled_from([0, 1, 2]), loops=st.integers(1, 120), sms=st.sampled_from([16, 37, 74, 148]),
       warps=st.sampled_from([8, 16]), nbr_rows=st.booleans(), nbr_cols=st.booleans())
def test_persist_plans_are_launchable(rows, cb, c, kind, loops, sms, warps, nbr_rows, nbr_cols):
    """Plans of the persistent kernel satisfy everything its launcher checks (persist_kernel.cu) and fit
    the exchange arena layout (xplan.h)."""
    from parallel_convolution_b200.parallel.fused import persist_arena_bytes
    cb = cb // c * c or c
    p = plan_persist(rows, cb, c, kind, loops, sms=sms, warps=warps,
                     min_nbr_rows=rows if nbr_rows else None, min_nbr_cols=cb // c if nbr_cols else None)
    if p is None:
        return
    assert p.W == warps and p.nty * p.ntx <= sms * (16 // warps) and p.nty * p.ntx <= 1024
    assert p.TH + 2 * p.k <= warps * p.R and p.TH >= p.k and 1 <= p.k <= loops
    assert p.HB % 16 == 0 and p.HB >= p.k * c and p.TWB + 2 * p.HB == 512 and p.TWB >= p.HB
    assert p.nty * p.TH >= rows > (p.nty - 1) * p.TH and p.ntx * p.TWB >= cb > (p.ntx - 1) * p.TWB
    if nbr_rows:
        assert rows - (p.nty - 1) * p.TH >= p.k
    if nbr_cols:
        assert cb % 16 == 0 and cb - (p.ntx - 1) * p.TWB >= p.HB
    assert persist_arena_bytes(p) > 0


@settings(max_examples=25, deadline=None)
@given(h=st.integers(20, 90), wpx=st.integers(40, 400), c=st.sampled_from([1, 3]), k=st.integers(1, 4),
       R=st.sampled_from([2, 3, 4]), warps=st.sampled_from([8, 16]), loops=st.integers(1, 9),
       seed=st.integers(0, 1000))
def test_emulated_exchange_random_geometry(h, wpx, c, k, R, warps, loops, seed):
    """Random single-rank geometries through the CPU replay of the persistent kernel's exchange."""
    if warps * R - 2 * k < k:
        return
    rng = np.random.default_rng(seed)
    img = rng.integers(0, 256, (h, wpx) if c == 1 else (h, wpx, c), dtype=np.uint8)
    f = Filter.builtin("gaussian")
    got = host.xplan_emulate(img, f.spec, loops, (1, 1), k, R, warps)
    if isinstance(got, int):
        assert got in (-2,)
        return
    assert np.array_equal(got, host.oracle(img, f.spec, loops))


@settings(max_examples=60, deadline=None)
@given(k=st.integers(1, 20), TH=st.integers(1, 220), nh=st.integers(1, 4), nty=st.integers(1, 30), ntx=st.integers(1, 30))
def test_arena_size_formula_matches_the_native_layout(k, TH, nh, nty, ntx):
    """parallel/fused.py:persist_arena_bytes mirrors csrc/common/xplan.h (xp_finish + xp_arena_bytes), which
    the native launcher evaluates for the same plan: flags area + two parities of one slot per CTA."""
    from parallel_convolution_b200.parallel.fused import persist_arena_bytes
    from parallel_convolution_b200.parallel.plan import StreamPlan
    HB = 16 * nh
    p = StreamPlan(k, 4, TH, 512 - 2 * HB, HB, ntx, nty, 1, 0.0, 8)
    nc = (512 - 2 * HB) // 16
    chunks = (2 * k * nc + 2 * TH * nh + 4 * k * nh + 7) // 8 * 8
    assert persist_arena_bytes(p) == 1024 * 16 * 4 + 2 * nty * ntx * chunks * 16


@settings(max_examples=80, deadline=None)
@given(height=st.integers(1, 100000), rb=st.integers(1, 200000), loops=st.integers(1, 500),
       mem=st.integers(1 << 16, 1 << 34), k=st.one_of(st.none(), st.integers(1, 64)))
def test_out_of_core_strip_plan(height, rb, loops, mem, k):
    """GPU out-of-core: strips cover the image, carry a halo no deeper than themselves, k never exceeds loops."""
    from parallel_convolution_b200.utils.out_of_core import plan_strips
    S, kp = plan_strips(height, rb, loops, mem, k)
    assert 1 <= kp <= loops and (k is None or kp <= k)
    assert 1 <= S <= height
    assert S >= min(height, 2 * kp)          # a strip is at least as deep as its two halos together


@settings(max_examples=20, deadline=None)
@given(gy=st.integers(1, 3), gx=st.integers(1, 3), tile_h=st.integers(12, 40), tile_chunks=st.integers(2, 24),
       c=st.sampled_from([1, 4]), k=st.integers(1, 3), R=st.sampled_from([2, 3, 4]), warps=st.sampled_from([8, 16]),
       loops=st.integers(1, 7), seed=st.integers(0, 1000))
def test_emulated_exchange_random_rank_grids(gy, gx, tile_h, tile_chunks, c, k, R, warps, loops, seed):
    """Random rank grids (up to 3x3: every one of the eight directions crosses a rank boundary somewhere,
    CTA coordinates wrap into the neighbouring rank) through the CPU replay of the pushed exchange."""
    tile_w_px = tile_chunks * 16 // c          # tile width = whole 16-byte chunks, as the launcher requires
    if tile_w_px < 1 or (tile_w_px * c) % 16:
        return
    h, w = gy * tile_h, gx * tile_w_px
    rng = np.random.default_rng(seed)
    img = rng.integers(0, 256, (h, w) if c == 1 else (h, w, c), dtype=np.uint8)
    f = Filter.builtin("gaussian")
    got = host.xplan_emulate(img, f.spec, loops, (gy, gx), k, R, warps)
    if isinstance(got, int):
        assert got in (-2, -4), got         # band too short for the halo / last CTA row or column too thin
        return
    assert np.array_equal(got, host.oracle(img, f.spec, loops))


@settings(max_examples=30, deadline=None)
@given(h=st.integers(6, 70), w=st.integers(6, 70), c=st.sampled_from([1, 3, 4]), gr=st.integers(1, 3), gc=st.integers(1, 3),
       k=st.integers(1, 7), loops=st.integers(0, 11), threads=st.integers(1, 2),
       filt=st.sampled_from(["gaussian", "box", "sharpen", "edge"]), seed=st.integers(0, 2**31))
def test_shared_image_ranks_random_geometry(h, w, c, gr, gc, k, loops, threads, filt, seed):
    """CPU ranks over the shared image (csrc/host/shm_ranks.cpp), ranks = threads of this process: any grid with
    remainder tiles, any depth up to the smallest tile, any iteration count -> the oracle's bytes."""
    import os
    import threading
    from parallel_convolution_b200.parallel.shm_backend import ShmTile
    if h < gr * 2 or w < gc * 2:
        return
    g = TileGrid.create(h, w, c, gr * gc, (gr, gc))
    k = min(k, g.min_tile_rows(), g.min_tile_cols()) if gr * gc > 1 else k
    f = Filter.builtin(filt)
    img = np.random.default_rng(seed).integers(0, 256, (h, w * c), dtype=np.uint8)
    path = f"/dev/shm/pconv_prop_{os.getpid()}_{seed}_{h}_{w}"
    seg = host.ShmSegment(path, h, w, c, gr * gc, create=True)
    os.unlink(path)
    tiles = [ShmTile(g, r, f, k, threads, None, 20_000.0, segment=seg) for r in range(gr * gc)]
    for t in tiles:
        tt = t.tile
        seg.reset(t.rank)
        t._view(0)[:] = img[tt.row0:tt.row0 + tt.rows, tt.col0 * c:(tt.col0 + tt.cols) * c]
    errs = []

    def work(t):
        try:
            t.iterate(loops)
        except Exception as e:  # noqa: BLE001
            errs.append(e)

    th = [threading.Thread(target=work, args=(t,)) for t in tiles]
    [x.start() for x in th]
    [x.join() for x in th]
    assert not errs, errs
    out = np.empty_like(img)
    for t in tiles:
        tt = t.tile
        out[tt.row0:tt.row0 + tt.rows, tt.col0 * c:(tt.col0 + tt.cols) * c] = t.result()
    shape = (h, w) if c == 1 else (h, w, c)
    want = host.oracle(img.reshape(shape), f.spec, loops).reshape(h, w * c)
    seg.close()
    assert np.array_equal(out, want)
