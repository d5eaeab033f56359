"""Host-side planner of the decoder launches (no GPU): page calls skip the decoder work outside the region
the reference's 9-case crop keeps (main.py:294-364) and choose their M-tile shapes for those regions.  A
missing work item would leave page pixels unlabelled, so the enumeration is checked against an independent
restatement: every low-res pixel that feeds a kept output pixel is covered by an item of the right parity."""
import ctypes as C

import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

from oracle import do_prediction as odp
from sbb_textline_detection_b200 import _lib


def plan(H, W, tile, margin, level, merged, full_grid):
    l = _lib.lib()
    bw, bh, n = C.c_int32(), C.c_int32(), C.c_int32()
    _lib.check(l.sbb_plan_decoder_tiles(H, W, tile, tile, margin, level, merged, full_grid, C.byref(bw), C.byref(bh),
                                        None, 0, C.byref(n)))
    items = np.zeros((n.value, 4), np.int32)
    _lib.check(l.sbb_plan_decoder_tiles(H, W, tile, tile, margin, level, merged, full_grid, C.byref(bw), C.byref(bh),
                                        items.ctypes.data_as(C.c_void_p), n.value, C.byref(n)))
    return bw.value, bh.value, items


def kept_boxes(H, W, tile, margin):
    """Per tile (reference loop order) the tile-coordinate box of the pixels whose write survives the stitch,
    from the oracle's literal replay of the reference loop."""
    m, nxf, nyf, tiles = odp.tile_grid(H, W, tile, tile, margin if margin >= 0 else None)
    assert len(tiles) < 255
    owner = odp.stitch_replay(H, W, tile, tile, m, nxf, nyf, tiles,
                              lambda t, i, j, x0, y0: np.full((tile, tile), t + 1, np.uint8))[:, :, 0].astype(int) - 1
    out = []
    for t, (_, _, x0, y0) in enumerate(tiles):
        ys, xs = np.nonzero(owner[y0:y0 + tile, x0:x0 + tile] == t)
        out.append(None if len(ys) == 0 else (xs.min(), ys.min(), xs.max(), ys.max()))
    return out


def needed(box, level, tile):
    """Region of decoder block `level`'s output needed for the kept level-5 box: each block reads its
    low-res input at +-1."""
    if box is None:
        return None
    x0, y0, x1, y1 = box
    size = tile
    for _ in range(5, level, -1):
        size //= 2
        x0, y0 = max(0, (x0 >> 1) - 1), max(0, (y0 >> 1) - 1)
        x1, y1 = min(size - 1, (x1 >> 1) + 1), min(size - 1, (y1 >> 1) + 1)
    return x0, y0, x1, y1


@pytest.mark.parametrize("H,W,tile,margin", [(2800, 2000, 448, -1), (4600, 3400, 672, -1), (600, 428, 96, -1),
                                             (1000, 900, 448, 20), (448, 448, 448, -1), (97, 131, 96, 3)])
@pytest.mark.parametrize("merged,full_grid", [(0, 0), (0, 1), (1, 0), (2, 0)])
def test_decoder_work_items_cover_the_kept_region(H, W, tile, margin, merged, full_grid):
    check_cover(H, W, tile, margin, merged, full_grid)


@settings(max_examples=40, deadline=None)
@given(tile=st.sampled_from([64, 96, 128]), dh=st.integers(0, 300), dw=st.integers(0, 300), margin=st.integers(0, 25),
       merged=st.integers(0, 2))
def test_decoder_work_items_cover_random_geometries(tile, dh, dw, margin, merged):
    """Ragged pages, clamped trailing tiles (several tiles at the same origin) and margins the reference never uses."""
    from hypothesis import assume
    wm = tile - 2 * margin
    assume(-(-(tile + dh) // wm) * -(-(tile + dw) // wm) < 255)   # kept_boxes labels tiles in a uint8 map
    check_cover(tile + dh, tile + dw, tile, margin, merged, 0)


def check_cover(H, W, tile, margin, merged, full_grid):
    boxes = kept_boxes(H, W, tile, margin)
    for level in ((5,) if merged == 1 else ((4,) if merged == 2 else (1, 2, 3, 4, 5))):
        G = tile >> (6 - level)  # half-resolution grid of the launch
        bw, bh, items = plan(H, W, tile, margin, level, merged, full_grid)
        assert 1 <= bw * bh <= 128
        assert (items[:, 2] >= 0).all() and (items[:, 3] >= 0).all() and (items[:, 2] < G).all() and (items[:, 3] < G).all()
        for t, box in enumerate(boxes):
            mine = items[items[:, 1] == t]
            r = needed(box, level, tile)
            if r is None:
                assert len(mine) == 0
                continue
            for par in ((0,) if merged == 1 else ((0, 1) if merged == 2 else (0, 1, 2, 3))):
                py, px = (par, -1) if merged == 2 else (par >> 1, par & 1)   # merged == 2: variant = row parity
                cover = np.zeros((G, G), bool)
                for _, _, X0, Y0 in mine[mine[:, 0] == par]:
                    cover[Y0:Y0 + bh, X0:X0 + bw] = True
                want = np.zeros((G, G), bool)
                if merged == 1:
                    want[r[1] >> 1:(r[3] >> 1) + 1, r[0] >> 1:(r[2] >> 1) + 1] = True
                elif merged == 2:   # both column parities: every low-res column with a needed output column
                    ys = [Y for Y in range(G) if r[1] <= 2 * Y + py <= r[3]]
                    if ys:
                        want[np.ix_(ys, list(range(r[0] >> 1, (r[2] >> 1) + 1)))] = True
                else:
                    ys = [Y for Y in range(G) if r[1] <= 2 * Y + py <= r[3]]
                    xs = [X for X in range(G) if r[0] <= 2 * X + px <= r[2]]
                    if ys and xs:
                        want[np.ix_(ys, xs)] = True
                assert not (want & ~cover).any(), (level, t, par)
                # no item lies entirely outside what is needed
                for _, _, X0, Y0 in mine[mine[:, 0] == par]:
                    assert want[Y0:Y0 + bh, X0:X0 + bw].any(), (level, t, par, X0, Y0)


def test_shapes_for_the_kept_regions_need_fewer_items_on_config2():
    """The point of choosing the M-tile shape for the kept regions (BASELINE config 2 geometry)."""
    for level in (2, 3):
        _, _, a = plan(2800, 2000, 448, -1, level, 0, 1)
        _, _, b = plan(2800, 2000, 448, -1, level, 0, 0)
        assert len(b) < 0.85 * len(a), (level, len(a), len(b))
    _, _, per_parity = plan(2800, 2000, 448, -1, 5, 0, 0)
    _, _, merged = plan(2800, 2000, 448, -1, 5, 1, 0)
    assert len(merged) <= 0.26 * len(per_parity)


def _chain_list(px, nA, nB, R):
    l = _lib.lib()
    n = C.c_int32()
    _lib.check(l.sbb_plan_chain_list(px, nA, nB, R, None, 0, C.byref(n)))
    buf = np.zeros((n.value, 2), np.int32)
    _lib.check(l.sbb_plan_chain_list(px, nA, nB, R, buf.ctypes.data, n.value, C.byref(n)))
    return buf


@settings(max_examples=60, deadline=None)
@given(st.integers(1, 40000), st.sampled_from([(4, 1), (8, 2), (16, 4), (2, 1), (3, 5)]), st.integers(1, 80))
def test_chained_work_list_order(px, n_tiles, R):
    """The work list of a chained launch (SBB_CHAIN=1) is what makes its per-M-tile counters safe: every work item
    exists exactly once, the two items of a pair position share variant and N tile and cover M tiles 2j / 2j+1, a
    reduce-conv item comes after ALL expand-conv items of its M tile (a persistent grid takes positions in order, so a
    consumer never waits for a producer that has not started), and until the expand conv runs out the list is made
    of rounds of R positions of one kind, reduce-conv rounds trailing their producers by three rounds."""
    nA, nB = n_tiles
    items = _chain_list(px, nA, nB, R)
    Mt = (px + 127) // 128
    Mp = (Mt + 1) // 2
    assert len(items) == 2 * Mp * (nA + nB)
    pairs = items.reshape(-1, 2, 2)
    assert np.array_equal(pairs[:, 0, 0], pairs[:, 1, 0])                      # same variant | N tile
    assert np.array_equal(pairs[:, 0, 1] + 128, pairs[:, 1, 1]) and not (pairs[:, 0, 1] % 256).any()
    variant, nt, j = pairs[:, 0, 0] & 255, pairs[:, 0, 0] >> 8, pairs[:, 0, 1] // 256
    seen = set(zip(variant.tolist(), nt.tolist(), j.tolist()))
    assert len(seen) == len(pairs) and seen == {(v, t, m) for v, n in ((0, nA), (1, nB)) for t in range(n) for m in range(Mp)}
    pos = np.arange(len(pairs))
    last_producer = np.full(Mp, -1)
    np.maximum.at(last_producer, j[variant == 0], pos[variant == 0])
    assert (pos[variant == 1] > last_producer[j[variant == 1]]).all()
    a_end = pos[variant == 0].max() + 1                                         # expand conv exhausted here
    full = (a_end // R) * R
    rounds = variant[:full].reshape(-1, R) if full else np.zeros((0, R), int)
    assert (rounds == rounds[:, :1]).all()                                      # homogeneous rounds
    early = (variant == 1) & (pos < a_end)
    assert (pos[early] - last_producer[j[early]] > 2 * R).all()                 # >= two whole rounds in between


def _simulate_chain(items, nA, R, flush_before_consumer=True):
    """Discrete model of conv_gemm_pair_kernel's chained-launch protocol: cluster c takes pair positions c, c+R, ...;
    its producer warp issues an item's loads in order and stalls at a reduce-conv item until the M pair's counter
    reached nA; its store threads post an expand-conv item's count ONE item later (deferred completion), or early at
    the top of a reduce-conv item (the flush), or after the loop.  Returns True when every item finishes."""
    pairs = items.reshape(-1, 2, 2)
    variant, j = pairs[:, 0, 0] & 255, pairs[:, 0, 1] // 256
    n = len(pairs)
    clusters = [list(range(c, n, R)) for c in range(min(R, n))]
    count = {}
    prod = [0] * len(clusters)            # items whose loads were issued
    epi = [0] * len(clusters)             # items finished by the epilogue
    flushed = [0] * len(clusters)         # items whose top-of-loop flush already ran
    pending = [None] * len(clusters)
    done_tail = [False] * len(clusters)
    progress = True
    while progress:
        progress = False
        for c, mine in enumerate(clusters):
            while prod[c] < len(mine) and (variant[mine[prod[c]]] == 0 or count.get(j[mine[prod[c]]], 0) >= nA):
                prod[c] += 1
                progress = True
            if epi[c] < len(mine):
                k = mine[epi[c]]
                if flushed[c] == epi[c]:                     # top of the item loop
                    flushed[c] += 1
                    if flush_before_consumer and variant[k] == 1 and pending[c] is not None:
                        count[pending[c]] = count.get(pending[c], 0) + 1
                        pending[c] = None
                    progress = True
                if epi[c] < prod[c]:                          # the item's operands arrived: MMAs, epilogue, stores
                    if pending[c] is not None:
                        count[pending[c]] = count.get(pending[c], 0) + 1
                    pending[c] = j[k] if variant[k] == 0 else None
                    epi[c] += 1
                    progress = True
            elif not done_tail[c]:
                done_tail[c] = True
                if pending[c] is not None:
                    count[pending[c]] = count.get(pending[c], 0) + 1
                    pending[c] = None
                progress = True
    return all(e == len(m) for e, m in zip(epi, clusters))


@settings(max_examples=40, deadline=None)
@given(st.integers(1, 30000), st.sampled_from([(4, 1), (8, 2), (16, 4), (2, 1), (3, 5)]), st.integers(1, 80))
def test_chained_launch_protocol_terminates(px, n_tiles, R):
    nA, nB = n_tiles
    assert _simulate_chain(_chain_list(px, nA, nB, R), nA, R)


def test_chained_launch_protocol_needs_the_flush():
    """The configuration that trapped on the GPU before the flush existed (12 tiles of 448 at stage 5: every reduce-conv
    item follows the last expand-conv items directly, and two clusters end up owing each other a count)."""
    items = _chain_list(12 * 196, 16, 4, 74)
    assert _simulate_chain(items, 16, 74)
    assert not _simulate_chain(items, 16, 74, flush_before_consumer=False)
