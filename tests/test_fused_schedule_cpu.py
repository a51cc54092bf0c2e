"""Host-only checks of the fused backward's static work-item schedule (mmgclip_b200/csrc/bwd_fused.cuh).

The kernel hands every CTA pair a fixed, key-ordered list of items: coefficient tiles ("A", type 0) of a block, and the
dA / dB K-slices ("B", types 1 / 2) that consume the block's coefficients.  B items of block s spin on doneA[s] == nA, A
items of block s spin on doneB[s - nbuf] == nB (the coefficient buffer they overwrite).  Nothing here needs a GPU:
`mmg_fused_bwd_schedule` enumerates the same cursor on the host, and the test replays all pairs against the counters to
prove that every item runs exactly once and that the spin-waits cannot dead-lock for any number of resident pairs.
"""
import ctypes

import pytest

from mmgclip_b200 import _lib

MAX_ITEMS = 1 << 16


def _schedule(rows, cols, D, n_owners, n_parts, part, pairs):
    lib = _lib.load()
    info = (ctypes.c_int * 8)()
    per_pair = []
    for pair in range(pairs):
        buf = (ctypes.c_int * (8 * MAX_ITEMS))()
        n = lib.mmg_fused_bwd_schedule(rows, cols, D, n_owners, n_parts, part, pairs, pair, buf, MAX_ITEMS, info)
        assert 0 <= n <= MAX_ITEMS
        per_pair.append([tuple(buf[8 * i:8 * i + 8]) for i in range(n)])
    keys = ("Rb", "Cb", "nbuf", "nA", "nB", "nblk", "kslI", "kslT")
    return per_pair, dict(zip(keys, info))


SHAPES = [
    # rows, cols, D, n_owners, n_parts, part
    (32768, 32768, 512, 1, 1, 0),      # BASELINE configs[1], one GPU
    (4096, 4096, 512, 1, 1, 0),        # configs[2]
    (4096, 32768, 512, 8, 1, 0),       # one rank of 8, rows sharded
    (4096, 32768, 512, 8, 2, 1),       # ... in two column parts (peer-reduce pipelining)
    (16384, 32768, 512, 2, 1, 0),
    (8192, 8192, 256, 1, 1, 0),
    (2048, 131072, 1024, 8, 1, 0),
]


@pytest.mark.parametrize("shape", SHAPES)
@pytest.mark.parametrize("pairs", [1, 3, 74])
def test_schedule_is_complete_and_deadlock_free(shape, pairs):
    rows, cols, D, n_owners, n_parts, part = shape
    per_pair, info = _schedule(rows, cols, D, n_owners, n_parts, part, pairs)
    if info["nblk"] == 0 or not any(per_pair):
        pytest.skip("shape not covered by the fused backward (block loop takes it)")
    nA, nB, nblk, nbuf = info["nA"], info["nB"], info["nblk"], info["nbuf"]

    # every item exactly once over all pairs
    seen = set()
    for items in per_pair:
        for it in items:
            key = it[:5]
            assert key not in seen, f"item {it} scheduled twice"
            seen.add(key)
    a_items = [k for k in seen if k[0] == 0]
    b_items = [k for k in seen if k[0] != 0]
    assert len(a_items) == nA * nblk
    assert len(b_items) == nB * nblk
    assert {k[1] for k in seen} == set(range(nblk))

    # the block order is a bijection between block indices and (row block, global column block) cells of the grid
    cell = {}
    for items in per_pair:
        for it in items:
            cell.setdefault(it[1], set()).add((it[7], it[6]))
    assert all(len(c) == 1 for c in cell.values())
    cells = {next(iter(c)) for c in cell.values()}
    assert len(cells) == nblk
    assert {c[0] for c in cells} == set(range(rows // info["Rb"]))

    # K-slices tile the contraction: per (type, block, tm, tn) the [kb0, kb0+nkb) ranges are disjoint and contiguous
    slices = {}
    for items in per_pair:
        for t, blk, tm, tn, kb0, nkb, _, _ in items:
            if t != 0:
                slices.setdefault((t, blk, tm, tn), []).append((kb0, nkb))
    for (t, blk, tm, tn), sl in slices.items():
        sl.sort()
        want = (info["Cb"] if t == 1 else info["Rb"]) // 64
        pos = 0
        for kb0, nkb in sl:
            assert kb0 == pos
            pos += nkb
        assert pos == want

    # replay: a pair runs its items in order and blocks on the counters exactly as the kernel does
    doneA = [0] * nblk
    doneB = [0] * nblk
    pos = [0] * pairs
    total = sum(len(x) for x in per_pair)
    done = 0
    while done < total:
        progressed = False
        for p in range(pairs):
            while pos[p] < len(per_pair[p]):
                t, blk = per_pair[p][pos[p]][:2]
                if t == 0:
                    if blk >= nbuf and doneB[blk - nbuf] != nB:
                        break
                    doneA[blk] += 1
                else:
                    if doneA[blk] != nA:
                        break
                    doneB[blk] += 1
                pos[p] += 1
                done += 1
                progressed = True
        assert progressed, f"dead-lock with {pairs} pairs at positions {pos}"
    assert doneA == [nA] * nblk and doneB == [nB] * nblk


def test_global_column_block_mapping_of_parts():
    """Two parts of an 8-owner launch cover disjoint halves of every owner's rows, together all column blocks."""
    rows, cols, D = 4096, 32768, 512
    cover = []
    for part in range(2):
        per_pair, info = _schedule(rows, cols, D, 8, 2, part, 4)
        if not any(per_pair):
            pytest.skip("shape not covered")
        cover.append({it[6] for items in per_pair for it in items})
    assert not (cover[0] & cover[1])
    assert cover[0] | cover[1] == set(range(cols // info["Cb"]))


def test_unsupported_shape_reports_zero():
    per_pair, _ = _schedule(100, 100, 64, 1, 1, 0, 2)
    assert per_pair == [[], []]


@pytest.mark.parametrize("sr,sc", [(1, 1), (2, 4), (4, 2), (8, 16), (3, 5)])
def test_block_order_super_tiles(sr, sc):
    """Any super-tile shape (non-dividing requests fall back to a dividing one) keeps the schedule complete and live."""
    lib = _lib.load()
    try:
        assert lib.mmg_tune(b"fused_sr", sr) == 0 and lib.mmg_tune(b"fused_sc", sc) == 0
        test_schedule_is_complete_and_deadlock_free((32768, 32768, 512, 1, 1, 0), 74)
        test_schedule_is_complete_and_deadlock_free((4096, 32768, 512, 8, 2, 1), 74)
    finally:
        assert lib.mmg_tune(b"reset", 0) == 0


def test_tune_rejects_unknown_keys():
    lib = _lib.load()
    assert lib.mmg_tune(b"no_such_knob", 1) != 0
    assert b"unknown key" in lib.mmg_last_error_string()


@pytest.mark.parametrize("shape", SHAPES)
@pytest.mark.parametrize("pairs", [1, 5, 74])
def test_stored_e_two_stream_replay(shape, pairs, defer=False):
    """Stored-E mode runs two independent streams per CTA pair: the transform warps walk the coefficient tiles (waiting
    for doneB of the block whose buffer they overwrite), the producer / MMA / epilogue warps walk the gradient slices
    (waiting for doneA of their block).  Both must run to completion."""
    rows, cols, D, n_owners, n_parts, part = shape
    per_pair, info = _schedule(rows, cols, D, n_owners, n_parts, part, pairs)
    if not any(per_pair):
        pytest.skip("shape not covered by the fused backward")
    nA, nB, nblk, nbuf = info["nA"], info["nB"], info["nblk"], info["nbuf"]
    if defer and nbuf < 3:
        pytest.skip("the launcher only selects the deferred kernel with at least three buffers")
    t_items = [[it for it in items if it[0] == 0] for items in per_pair]
    m_items = [[it for it in items if it[0] != 0] for items in per_pair]
    doneA, doneB = [0] * nblk, [0] * nblk
    tpos, mpos = [0] * pairs, [0] * pairs
    pending = [None] * pairs
    remaining = sum(len(x) for x in t_items) + sum(len(x) for x in m_items)
    while remaining:
        progressed = False
        for p in range(pairs):
            while tpos[p] < len(t_items[p]):
                blk = t_items[p][tpos[p]][1]
                if blk >= nbuf and doneB[blk - nbuf] != nB:
                    break
                if defer:
                    if pending[p] is not None:
                        doneA[pending[p]] += 1
                    pending[p] = blk
                    if tpos[p] + 1 == len(t_items[p]):
                        doneA[blk] += 1
                        pending[p] = None
                else:
                    doneA[blk] += 1
                tpos[p] += 1
                remaining -= 1
                progressed = True
            while mpos[p] < len(m_items[p]):
                blk = m_items[p][mpos[p]][1]
                if doneA[blk] != nA:
                    break
                doneB[blk] += 1
                mpos[p] += 1
                remaining -= 1
                progressed = True
        assert progressed, f"dead-lock (defer={defer}) with {pairs} pairs: transform at {tpos}, slices at {mpos}"
    assert doneA == [nA] * nblk and doneB == [nB] * nblk


def _needed_tiles(it, info):
    """Coefficient tiles (tm, tn) whose scratch data a gradient slice reads."""
    t, blk, tm, tn, kb0, nkb = it[:6]
    lo, hi = kb0 * 64 // 256, ((kb0 + nkb) * 64 - 1) // 256
    if t == 1:   # dA slice of row panel tm: K runs over the block's columns
        return [(tm, c) for c in range(lo, hi + 1)]
    return [(r, tm) for r in range(lo, hi + 1)]   # dB slice of column panel tm: K runs over the block's rows


@pytest.mark.parametrize("shape", SHAPES[:5])
@pytest.mark.parametrize("pairs", [3, 74])
@pytest.mark.parametrize("mode", ["recompute", "stored"])
def test_dependency_counters_cover_the_data_each_slice_reads(shape, pairs, mode, panel=False):
    """Replays the kernel's own counter arithmetic (block counters) and checks, at the moment a gradient slice is allowed to start, that every coefficient tile it reads is complete -- and
    that everything runs to completion."""
    rows, cols, D, n_owners, n_parts, part = shape
    per_pair, info = _schedule(rows, cols, D, n_owners, n_parts, part, pairs)
    if not any(per_pair):
        pytest.skip("shape not covered by the fused backward")
    nA, nB, nblk, nbuf = info["nA"], info["nB"], info["nblk"], info["nbuf"]
    tAm, tAn = info["Rb"] // 256, info["Cb"] // 256
    assert nA == tAm * tAn
    done_tiles = set()                     # (blk, tm, tn) published
    blkA = [0] * nblk
    rowA = [[0] * tAm for _ in range(nblk)]
    colA = [[0] * tAn for _ in range(nblk)]
    doneB = [0] * nblk

    def publish(it):
        _, blk, tm, tn = it[:4]
        done_tiles.add((blk, tm, tn))
        blkA[blk] += 1
        rowA[blk][tm] += 1
        colA[blk][tn] += 1

    def b_may_start(it):
        t, blk, tm = it[0], it[1], it[2]
        if not panel:
            return blkA[blk] == nA
        return rowA[blk][tm] == tAn if t == 1 else colA[blk][tm] == tAm

    def a_may_start(it):
        blk = it[1]
        return blk < nbuf or doneB[blk - nbuf] == nB

    if mode == "stored":
        streams = [[it for it in items if it[0] == 0] for items in per_pair] + \
                  [[it for it in items if it[0] != 0] for items in per_pair]
    else:
        streams = [list(items) for items in per_pair]
    pos = [0] * len(streams)
    pending = [None] * len(streams)        # recompute mode: tile published when the stream reaches its next item
    remaining = sum(len(x) for x in streams)
    while remaining:
        progressed = False
        for si, st in enumerate(streams):
            while pos[si] < len(st):
                it = st[pos[si]]
                if mode == "recompute" and pending[si] is not None:
                    publish(pending[si])
                    pending[si] = None
                    progressed = True
                if it[0] == 0:
                    if not a_may_start(it):
                        break
                    if mode == "recompute":
                        pending[si] = it
                    else:
                        publish(it)
                else:
                    if not b_may_start(it):
                        break
                    for tm, tn in _needed_tiles(it, info):
                        assert (it[1], tm, tn) in done_tiles, f"slice {it} started before tile ({tm},{tn}) of its block"
                    doneB[it[1]] += 1
                pos[si] += 1
                remaining -= 1
                progressed = True
            if mode == "recompute" and pos[si] == len(st) and pending[si] is not None:
                publish(pending[si])
                pending[si] = None
                progressed = True
        assert progressed, f"dead-lock (mode={mode}, panel={panel}) at {pos}"
    assert len(done_tiles) == nA * nblk and doneB == [nB] * nblk
