"""Differential tests of the oracle itself, in the shape of the reference's proptests:
Block-WAND (pruned, reference-faithful) against the exhaustive canonical top-k
  src/query/boolean_query/block_wand_union.rs:427-504   (1 and 2+ scorers, docs repeated x64)
  src/query/boolean_query/block_wand_intersection.rs:335-424 (2 and 3 scorers)
plus codec round trips (src/postings/compression/mod.rs:276-351) and search_block vs a linear
scan (src/postings/block_search.rs:163-179)."""
import ctypes as C

import numpy as np
import pytest

from oracle import tq_oracle as O
from tantivy_b200._abi import TQ_OP_AND, TQ_OP_OR, TQ_OP_TERM, QueryBatch
from tests.helpers import OracleSegment, hits, make_query

REPEAT = 64


def _expanded(rng, num_scorers, max_doc):
    lists = []
    for _ in range(num_scorers):
        df = int(rng.integers(1, max_doc + 1))
        docs = np.sort(rng.choice(max_doc, size=df, replace=False))
        tfs = rng.integers(1, 100, size=df)
        ed = (docs[:, None] * REPEAT + np.arange(REPEAT)[None, :]).reshape(-1)
        et = np.ones((df, REPEAT), dtype=np.int64)
        et[:, 0] = tfs
        lists.append((ed.astype(np.uint32), et.reshape(-1).astype(np.uint32)))
    fieldnorms = np.repeat(rng.integers(2, 1000, size=max_doc), REPEAT)
    return lists, fieldnorms


def _same_topk(a, b):
    assert len(a) == len(b)
    for (sa, ga, da), (sb, gb, db) in zip(a, b):
        if (ga, da) != (gb, db):
            # the pruned path may swap documents whose scores are equal up to summation order
            assert abs(sa - sb) <= 1e-5 * max(abs(sa), abs(sb))
        else:
            assert abs(sa - sb) <= 1e-5 * max(abs(sa), abs(sb))
    return sum((ga, da) != (gb, db) for (_, ga, da), (_, gb, db) in zip(a, b))


@pytest.mark.parametrize("op,num_scorers", [(TQ_OP_TERM, 1), (TQ_OP_OR, 2), (TQ_OP_OR, 3), (TQ_OP_AND, 2), (TQ_OP_AND, 3)])
def test_block_wand_matches_exhaustive(op, num_scorers):
    rng = np.random.default_rng(100 * op + num_scorers)
    swapped = 0
    total = 0
    for case in range(60):
        max_doc = int(rng.integers(1, 100))
        lists, fieldnorms = _expanded(rng, num_scorers, max_doc)
        seg = OracleSegment(lists, fieldnorms)
        ix = O.OracleIndex()
        seg.register(ix)
        for k in (1, 2, 3, 10):
            batch = QueryBatch([make_query(op, [seg], list(range(num_scorers)), k)])
            exhaustive = hits(ix.search_batch(batch, mode=0))
            pruned = hits(ix.search_batch(batch, mode=1))
            swapped += _same_topk(exhaustive, pruned)
            total += len(exhaustive)
    # single term and AND sum in a fixed order in both paths: must agree exactly
    if op != TQ_OP_OR:
        assert swapped == 0
    assert swapped <= 0.02 * max(total, 1)


def test_multi_segment_merge_matches_canonical():
    rng = np.random.default_rng(5)
    for case in range(20):
        segs = []
        for so in range(3):
            max_doc = int(rng.integers(200, 3000))
            lists = []
            for t in range(3):
                df = int(rng.integers(1, max_doc))
                docs = np.sort(rng.choice(max_doc, size=df, replace=False)).astype(np.uint32)
                lists.append((docs, rng.integers(1, 6, size=df).astype(np.uint32)))
            segs.append(OracleSegment(lists, rng.integers(1, 300, size=max_doc), segment_ord=so))
        ix = O.OracleIndex()
        for s in segs:
            s.register(ix)
        for op, terms in [(TQ_OP_TERM, [1]), (TQ_OP_AND, [0, 1, 2]), (TQ_OP_OR, [0, 2])]:
            batch = QueryBatch([make_query(op, segs, terms, 25)])
            a = hits(ix.search_batch(batch, mode=0))
            b = hits(ix.search_batch(batch, mode=1))
            _same_topk(a, b)
            # canonical ordering: score desc, segment asc, doc asc
            assert a == sorted(a, key=lambda h: (-h[0], h[1], h[2]))


def test_codec_roundtrip_all_bit_widths():
    L = O.lib()
    rng = np.random.default_rng(1)
    out = np.zeros(640, dtype=np.uint8)
    written = C.c_size_t()
    dec = np.zeros(128, dtype=np.uint32)
    dec2 = np.zeros(128, dtype=np.uint32)
    for b in range(0, 33):
        hi = (1 << b) - 1
        vals = rng.integers(0, hi + 1, size=128, dtype=np.uint64).astype(np.uint32)
        if b:
            vals[int(rng.integers(0, 128))] = hi
        nb = L.tqo_compress_block_unsorted(O.ptr(vals, O.u32p), 0, O.ptr(out, O.u8p), C.byref(written))
        assert nb == b and written.value == 16 * b
        padded = np.concatenate([out[:16 * b], np.zeros(32, np.uint8)])
        L.tqo_bp4x_unpack(O.ptr(padded, O.u8p), b, O.ptr(dec, O.u32p))
        L.tqo_bp4x_unpack_scalar(O.ptr(padded, O.u8p), b, O.ptr(dec2, O.u32p))
        assert (dec == vals).all() and (dec2 == vals).all()
        if b < 32 and b > 0:
            # strictly sorted with that many delta bits, both offset == 0 (None) and offset > 0
            for offset in (0, 12345):
                gaps = rng.integers(0, min(hi, 2 ** 20) + 1, size=128, dtype=np.uint64)
                gaps[3] = min(hi, 2 ** 20)
                start = 0 if offset == 0 else offset + 1
                docs = (start + np.cumsum(gaps + 1) - (gaps[0] + 1) + gaps[0]).astype(np.uint64)
                if docs[-1] >= 2 ** 31 - 1:
                    continue
                docs = docs.astype(np.uint32)
                nb = L.tqo_compress_block_sorted(O.ptr(docs, O.u32p), offset, O.ptr(out, O.u8p), C.byref(written))
                L.tqo_uncompress_block_sorted(O.ptr(out, O.u8p), offset, nb, 1, O.ptr(dec, O.u32p))
                assert (dec == docs).all()


def test_codec_sorted_wide_doc_widths():
    """Doc bit widths 21..31 (one large strict gap per block; doc ids stay below TERMINATED), offset 0 and > 0."""
    L = O.lib()
    rng = np.random.default_rng(5)
    out = np.zeros(640, dtype=np.uint8)
    written = C.c_size_t()
    dec = np.zeros(128, dtype=np.uint32)
    for b in range(21, 32):
        for offset in (0, 777):
            gaps = rng.integers(0, 50, size=128, dtype=np.uint64)  # stored value = gap - 1 (strict delta)
            gaps[int(rng.integers(0, 128))] = 2 ** (b - 1) + int(rng.integers(0, 2 ** (b - 1) - 200)) if b < 31 else 2 ** 30
            first = gaps[0] if offset == 0 else offset + 1 + gaps[0]
            docs = first + np.concatenate([[0], np.cumsum(gaps[1:] + 1)])
            assert docs[-1] < 2 ** 31 - 1
            docs = docs.astype(np.uint32)
            nb = L.tqo_compress_block_sorted(O.ptr(docs, O.u32p), offset, O.ptr(out, O.u8p), C.byref(written))
            assert (nb & 31) == b and written.value == 16 * b
            L.tqo_uncompress_block_sorted(O.ptr(out, O.u8p), offset, nb, 1, O.ptr(dec, O.u32p))
            assert (dec == docs).all()


def test_search_block_matches_linear_scan():
    L = O.lib()
    rng = np.random.default_rng(2)
    for _ in range(200):
        n = int(rng.integers(1, 129))
        arr = np.sort(rng.choice(5000, size=n, replace=False)).astype(np.uint32)
        padded = np.concatenate([arr, np.full(128 - n, 0x7FFFFFFF, np.uint32)])
        for target in list(arr[:5]) + [0, 1, int(arr[-1]), int(arr[-1]) + 1, 4999]:
            if target > int(padded[-1]):
                continue
            assert L.tqo_search_block(O.ptr(padded, O.u32p), int(target)) == int((padded < target).sum())


def test_basic_requested_on_freq_field_scores_with_tf_one():
    """FreqReadingOption::SkipFreq (block_segment_postings.rs:97-140): the Basic-requested scores of a
    WithFreqs list equal the scores of the same docs written with every tf = 1."""
    rng = np.random.default_rng(11)
    max_doc = 3000
    fieldnorms = rng.integers(1, 300, size=max_doc)
    docs = np.sort(rng.choice(max_doc, size=700, replace=False)).astype(np.uint32)  # 5 blocks + a VInt tail
    tfs = rng.integers(1, 9, size=len(docs)).astype(np.uint32)
    with_tf = OracleSegment([(docs, tfs)], fieldnorms, segment_ord=0)
    ones = OracleSegment([(docs, np.ones_like(tfs))], fieldnorms, segment_ord=0)
    out = []
    for seg, flags in ((with_tf, [1]), (ones, None), (with_tf, None)):
        ix = O.OracleIndex()
        seg.register(ix)
        q = make_query(TQ_OP_TERM, [seg], [0], 50)
        if flags:
            q["term_flags"] = flags
        out.append([hits(ix.search_batch(QueryBatch([q]), mode=m)) for m in (0, 1)])
    assert out[0][0] == out[1][0] and out[0][1] == out[1][1]
    assert out[0][0] != out[2][0]


def test_initial_threshold_exhaustive_vs_pruned():
    """for_each_pruning(threshold, ..) (weight.rs:123-132): with an initial threshold both oracle paths return the
    docs scoring strictly above it."""
    rng = np.random.default_rng(12)
    max_doc = 4000
    lists = []
    for p in (0.4, 0.1, 0.02):
        docs = np.nonzero(rng.random(max_doc) < p)[0].astype(np.uint32)
        lists.append((docs, rng.integers(1, 6, size=len(docs)).astype(np.uint32)))
    seg = OracleSegment(lists, rng.integers(1, 300, size=max_doc))
    ix = O.OracleIndex()
    seg.register(ix)
    for op, terms in [(TQ_OP_TERM, [1]), (TQ_OP_AND, [0, 1]), (TQ_OP_OR, [0, 1, 2])]:
        q = make_query(op, [seg], terms, 30)
        full = hits(ix.search_batch(QueryBatch([q]), mode=0))
        thr = full[12][0]
        q["threshold"] = float(thr)
        a = hits(ix.search_batch(QueryBatch([q]), mode=0))
        b = hits(ix.search_batch(QueryBatch([q]), mode=1))
        assert a == [h for h in full if h[0] > thr]
        assert [(g, d) for _, g, d in a] == [(g, d) for _, g, d in b]


def test_count_matches_set_arithmetic():
    """The oracle's Count collector against numpy set arithmetic on the posting lists."""
    rng = np.random.default_rng(13)
    max_doc = 5000
    lists = []
    for p in (0.4, 0.1, 0.02):
        docs = np.nonzero(rng.random(max_doc) < p)[0].astype(np.uint32)
        lists.append((docs, np.ones(len(docs), np.uint32)))
    alive_bits = rng.random(max_doc) > 0.3
    seg = OracleSegment(lists, rng.integers(1, 300, size=max_doc), alive=np.packbits(alive_bits, bitorder="little"))
    ix = O.OracleIndex()
    seg.register(ix)
    sets = [set(int(d) for d in docs if alive_bits[d]) for docs, _ in lists]
    qb = QueryBatch([make_query(TQ_OP_TERM, [seg], [0], 1), make_query(TQ_OP_AND, [seg], [0, 1], 1), make_query(TQ_OP_AND, [seg], [0, 1, 2], 1),
                     make_query(TQ_OP_OR, [seg], [1, 2], 1), make_query(TQ_OP_OR, [seg], [0, 1, 2], 1)])
    want = [len(sets[0]), len(sets[0] & sets[1]), len(sets[0] & sets[1] & sets[2]), len(sets[1] | sets[2]), len(sets[0] | sets[1] | sets[2])]
    assert [int(x) for x in ix.count_batch(qb)] == want
