"""Pins the CPU oracle against every known-answer test the reference holds for the hot path
(SURVEY.md §8c).  Each test cites the reference test it reproduces (paths relative to
/root/reference).  Runs on CPU."""
import math

import numpy as np
import pytest

from oracle import tq_oracle as O
from tantivy_b200._abi import (TERMINATED, TQ_OP_AND, TQ_OP_OR, TQ_OP_TERM, TQ_RECORD_BASIC, TQ_RECORD_FREQS,
                               TQ_RECORD_FREQS_POSITIONS, QueryBatch)
from tests.helpers import GOLDEN, OracleSegment, f32, hits, make_query

REL = 2e-7  # golden constants are printed with 8 significant digits (f32 shortest repr)


def near(a, b, rel=REL):
    return abs(float(a) - float(b)) <= rel * max(abs(float(a)), abs(float(b)), 1e-30)


# ---- fieldnorm code ------------------------------------------------------------------------------
def test_fieldnorm_table_matches_reference():
    # src/fieldnorm/code.rs:13 FIELD_NORMS_TABLE, tests `test_fieldnorm_byte`
    table = GOLDEN["field_norms_table"]
    assert [O.id_to_fieldnorm(i) for i in range(256)] == table
    for i, v in enumerate(table):
        assert O.fieldnorm_to_id(v) == i
    # code.rs tests: fieldnorm_to_id(41)=40, (42)=41, u32::MAX -> 255
    assert O.fieldnorm_to_id(41) == 40 and O.fieldnorm_to_id(42) == 41
    assert O.fieldnorm_to_id(0xFFFFFFFF) == 255
    # src/fieldnorm/reader.rs:168-193: 1_000_000 -> 983_064 ; constant(300) -> id 72 -> 280
    assert O.id_to_fieldnorm(O.fieldnorm_to_id(1_000_000)) == 983_064
    assert O.fieldnorm_to_id(300) == 72 and O.id_to_fieldnorm(72) == 280


# ---- BM25 ------------------------------------------------------------------------------------------
def test_idf():
    # src/query/bm25.rs:236-239
    assert near(O.bm25_idf(1, 2), math.log(2.0), 1e-6)


def _scorer(doc_and_tfs, fieldnorms, n, N, avg):
    """TermScorer::create_for_test + Bm25Weight::for_one_term(n, N, avg)."""
    docs = [d for d, _ in doc_and_tfs]
    tfs = [t for _, t in doc_and_tfs]
    seg = OracleSegment([(docs, tfs)], fieldnorms)
    ix = O.OracleIndex()
    seg.register(ix)
    w = O.bm25_weight(n, N)
    return ix, seg, ix.term_scorer(seg.term_seg(0), w, avg), w


def test_term_scorer_max_score():
    # src/query/term_query/term_scorer.rs:168-191
    ix, seg, ts, _ = _scorer([(2, 3), (3, 12), (7, 8)], [0, 0, 10, 12, 0, 0, 0, 100], 3, 6, 10.0)
    assert near(ts.max_score(), 1.3990127)
    assert ts.doc() == 2 and ts.term_freq() == 3
    assert near(ts.block_max_score(), 1.3676447)
    assert near(ts.score(), 1.0892314)
    assert ts.advance() == 3 and ts.doc() == 3 and ts.term_freq() == 12
    assert near(ts.score(), 1.3676447)
    assert ts.advance() == 7 and ts.doc() == 7 and ts.term_freq() == 8
    assert near(ts.score(), 0.72015285)
    assert ts.advance() == TERMINATED


def test_term_scorer_shallow_advance():
    # src/query/term_query/term_scorer.rs:193-209
    doc_and_tfs = [(i * 10, 1 + (i * 10) % 3) for i in range(300)]
    ix, seg, ts, _ = _scorer(doc_and_tfs, [10] * 3000, 300, 1024, 10.0)
    assert ts.doc() == 0
    ts.seek_block(1289)
    assert ts.doc() == 0
    ts.seek(1289)
    assert ts.doc() == 1290


def test_block_wand_block_max():
    # src/query/term_query/term_scorer.rs:254-278 (test_block_wand)
    doc_tfs = [(d, 1) for d in range(128)] + [(d, 2 if d == 200 else 1) for d in range(128, 256)]
    doc_tfs += [(256, 1), (257, 3), (258, 1)]
    ix, seg, ts, _ = _scorer(doc_tfs, [20] * 300, 10, 129, 20.0)
    assert near(ts.block_max_score(), 2.5161593)
    ts.seek_block(135)
    assert near(ts.block_max_score(), 3.4597192)
    ts.seek_block(256)
    assert near(ts.block_max_score(), 5.2971773)  # block not loaded -> global max_score()
    assert ts.seek(256) == 256
    assert near(ts.block_max_score(), 3.9539647)


def test_term_scorer_block_max_equals_recomputed():
    # src/query/term_query/term_scorer.rs:211-252 (proptest) on seeded random inputs
    rng = np.random.default_rng(7)
    for _ in range(40):
        n = int(rng.integers(80, 300))
        tfs = rng.integers(1, 10, n)
        extra = rng.integers(0, 100, n)
        fieldnorms = (tfs + extra).astype(np.uint32)
        avg = f32(f32(int(fieldnorms.sum())) / f32(n))
        ix, seg, ts, _ = _scorer(list(zip(range(n), tfs.tolist())), fieldnorms, n, n * 10, float(avg))
        doc = 0
        while doc < n:
            bm = ts.block_max_score()
            computed = 0.0
            for d in range(doc, min(doc + 128, n)):
                assert ts.doc() == d
                computed = max(computed, ts.score())
                ts.advance()
            assert abs(computed - bm) <= 0.0005
            doc += 128


# ---- end-to-end known answers ---------------------------------------------------------------------
def _search(segments, op, terms, k, mode):
    ix = O.OracleIndex()
    for s in segments:
        s.register(ix)
    batch = QueryBatch([make_query(op, segments, terms, k)])
    return hits(ix.search_batch(batch, mode=mode))


@pytest.mark.parametrize("mode", [0, 1])
def test_droopy_tax(mode):
    # src/collector/top_score_collector.rs:718-729,838-857 — query "droopy tax" (default OR)
    # doc0 "Hello happy tax payer." doc1 "Droopy says hello happy tax payer" doc2 "I like Droopy"
    seg = OracleSegment([([1, 2], [1, 1]), ([0, 1], [1, 1])], [4, 6, 3], record_option=TQ_RECORD_FREQS_POSITIONS)
    res = _search([seg], TQ_OP_OR, [0, 1], 4, mode)
    assert [(s, d) for _, s, d in res] == [(0, 1), (0, 2), (0, 0)]
    for got, want in zip(res, [0.81221175, 0.5376842, 0.48527452]):
        assert near(got[0], want, 1e-6)


@pytest.mark.parametrize("mode", [0, 1])
def test_term_weight(mode):
    # src/query/term_query/mod.rs:82-129 — field "left": doc0 18 tokens (left2 x2, left1 x1), doc1 "left4 left1"
    seg = OracleSegment([([0], [2]), ([0, 1], [1, 1])], [18, 2], record_option=TQ_RECORD_FREQS_POSITIONS)
    res = _search([seg], TQ_OP_TERM, [0], 2, mode)
    assert len(res) == 1 and near(res[0][0], 0.77802235, 1e-6)
    res = _search([seg], TQ_OP_TERM, [1], 2, mode)
    assert len(res) == 2 and near(res[0][0], 0.27101856, 1e-6) and near(res[1][0], 0.13736556, 1e-6)
    res = _search([seg], TQ_OP_OR, [0, 1], 2, mode)
    assert len(res) == 2 and near(res[0][0], 0.9153879, 1e-6) and near(res[1][0], 0.27101856, 1e-6)


def test_term_query_no_freq():
    # src/query/term_query/mod.rs:21-44 — one doc "a" in a STRING field (Basic), score 0.28768212
    seg = OracleSegment([([0], None)], [1], record_option=TQ_RECORD_BASIC)
    res = _search([seg], TQ_OP_TERM, [0], 1, 0)
    assert res[0][2] == 0 and near(res[0][0], 0.28768212, 1e-6)


@pytest.mark.parametrize("mode", [0, 1])
def test_boolean_query_with_weight(mode):
    # src/query/boolean_query/mod.rs:221-259 — docs "a b c","a c","b c"; a OR b; doc0 0.84163445, boost 2 -> 1.6832689
    seg = OracleSegment([([0, 1], [1, 1]), ([0, 2], [1, 1])], [3, 2, 2], record_option=TQ_RECORD_FREQS_POSITIONS)
    res = _search([seg], TQ_OP_OR, [0, 1], 3, mode)
    by_doc = {d: s for s, _, d in res}
    assert near(by_doc[0], 0.84163445, 1e-6)
    ix = O.OracleIndex()
    seg.register(ix)
    res2 = hits(ix.search_batch(QueryBatch([make_query(TQ_OP_OR, [seg], [0, 1], 3, boost=2.0)]), mode=mode))
    assert near({d: s for s, _, d in res2}[0], 1.6832689, 1e-6)


@pytest.mark.parametrize("mode", [0, 1])
def test_intersection_score(mode):
    # src/query/boolean_query/mod.rs:262-291 — docs "a b c","a c","b c","a b c d","d"; +a +b -> 0.977973, 0.84699446
    seg = OracleSegment([([0, 1, 3], [1, 1, 1]), ([0, 2, 3], [1, 1, 1])], [3, 2, 2, 4, 1], record_option=TQ_RECORD_FREQS_POSITIONS)
    res = _search([seg], TQ_OP_AND, [0, 1], 10, mode)
    by_doc = {d: s for s, _, d in res}
    assert sorted(by_doc) == [0, 3]
    assert near(by_doc[0], 0.977973, 1e-6) and near(by_doc[3], 0.84699446, 1e-6)


# ---- skip list / codec known answers ----------------------------------------------------------------
def test_skip_encodings():
    L = O.lib()
    # src/postings/skip.rs:316-330
    for tf in range(255):
        assert L.tqo_encode_block_wand_max_tf(tf) == tf and L.tqo_decode_block_wand_max_tf(tf) == tf
    for tf in (255, 256, 1_000_000, 0xFFFFFFFF):
        assert L.tqo_encode_block_wand_max_tf(tf) == 255
    assert L.tqo_decode_block_wand_max_tf(255) == 0xFFFFFFFF
    # src/postings/skip.rs:451-462
    assert L.tqo_encode_bitwidth(0b10, 1) == 0b01000010 and L.tqo_encode_bitwidth(0b10, 0) == 0b00000010


@pytest.mark.parametrize("record_option,rec", [(TQ_RECORD_BASIC, 5), (TQ_RECORD_FREQS, 8), (TQ_RECORD_FREQS_POSITIONS, 12)])
def test_skip_record_layout(record_option, rec):
    # skip.rs:205-253 record sizes; serializer.rs:471-476 VInt(skip_len) + records precede the blocks
    n = 128 * 3 + 5
    docs = np.arange(n, dtype=np.uint32) * 3
    tfs = None if record_option == TQ_RECORD_BASIC else (1 + (np.arange(n) % 4)).astype(np.uint32)
    seg = OracleSegment([(docs, tfs)], [10] * (3 * n), record_option=record_option)
    body = seg.body[8:]
    assert body[0] == (3 * rec) | 0x80  # VInt(skip_len), single byte with stop bit
    for b in range(3):
        r = body[1 + b * rec: 1 + (b + 1) * rec]
        assert int.from_bytes(bytes(r[:4]), "little") == int(docs[128 * (b + 1) - 1])
        assert r[4] == (2 | 0x40)  # strict delta: gaps of 3 -> value 2 -> 2 bits, delta-1 flag
        if record_option != TQ_RECORD_BASIC:
            assert r[5] == 2  # tf-1 in 0..3
    assert int.from_bytes(bytes(seg.body[:8]), "little") == seg.total_num_tokens


def test_compressed_sizes_and_vint():
    L = O.lib()
    import ctypes as C
    # compression/mod.rs:283-285,320-322: sorted 0..128 -> 0 bits?? (strict delta of i vs offset 0/None = 0) ; sizes = 16*b
    vals = np.arange(128, dtype=np.uint32)
    out = np.zeros(640, dtype=np.uint8)
    written = C.c_size_t()
    nb = L.tqo_compress_block_sorted(O.ptr(vals, O.u32p), 0, O.ptr(out, O.u8p), C.byref(written))
    assert nb == 0 and written.value == 0
    vals = (np.arange(128, dtype=np.uint32) * 7) % 12
    nb = L.tqo_compress_block_unsorted(O.ptr(vals, O.u32p), 0, O.ptr(out, O.u8p), C.byref(written))
    assert nb == 4 and written.value == 64
    # compression/mod.rs:359-376 test_encode_vint: 123 values <= 154 bytes, padding kept
    inp = np.array([4 + i * 7 // 2 for i in range(123)], dtype=np.uint32)
    for offset in (0, 1, 2):
        n = L.tqo_vint_compress_sorted(O.ptr(inp, O.u32p), 123, offset, O.ptr(out, O.u8p))
        assert n <= 154
        dec = np.zeros(128, dtype=np.uint32)
        consumed = L.tqo_vint_uncompress_sorted(O.ptr(out, O.u8p), 123, offset, 234_234_345, O.ptr(dec, O.u32p))
        assert consumed == n and (dec[:123] == inp).all() and (dec[123:] == 234_234_345).all()


def test_compat_fixture_v7_postings():
    # tests/compat_tests_data/index_v7/*.idx (src/compat_tests.rs:39-56): field 0 body is
    # u64 total_num_tokens=1 followed by a single 1-doc VInt posting list (doc 0, tf 1).
    idx = bytes.fromhex(GOLDEN["compat"]["index_v7"]["idx"])
    body0 = np.frombuffer(idx[:10], dtype=np.uint8)
    assert int.from_bytes(idx[:8], "little") == 1
    ix = O.OracleIndex()
    ix.segment_register(0, 0, 1, TQ_RECORD_FREQS_POSITIONS, body0, np.array([1], dtype=np.uint8))
    docs, tfs = ix.decode_postings((0, 0, 0, 1, 0, 2))
    assert docs.tolist() == [0] and tfs.tolist() == [1]
    # our own writer reproduces those bytes
    seg = OracleSegment([([0], [1])], [1], record_option=TQ_RECORD_FREQS_POSITIONS)
    assert bytes(seg.body) == idx[:10]


# ---- collectors ----------------------------------------------------------------------------------------
def test_top_n_heap_semantics():
    # src/collector/sort_key/sort_by_score.rs:171-247
    thr, res = O.top_n_heap([1.0, 2.0], [0, 1], 0)
    assert res == []
    thr, res = O.top_n_heap([1.0, 3.0, 2.0], [0, 1, 2], 2)
    assert res == [(3.0, 1), (2.0, 2)]
    thr, res = O.top_n_heap([1.0, 3.0, 2.0, 4.0], [0, 1, 2, 3], 2)
    assert thr.tolist() == [-math.inf, 1.0, 2.0, 3.0]
    thr, res = O.top_n_heap([5.0, 5.0, 5.0], [0, 1, 2], 2)  # tie: lower doc wins, equal rejected
    assert sorted(d for _, d in res) == [0, 1]
    thr, res = O.top_n_heap([1.0, 0.5, 2.0], [0, 1, 2], 1)
    assert res == [(2.0, 2)] and thr.tolist() == [1.0, 1.0, 2.0]
    thr, res = O.top_n_heap([3.0, 1.0, 2.0], [0, 1, 2], 5)
    assert res == [(3.0, 0), (2.0, 2), (1.0, 1)] and all(t == -math.inf for t in thr)


def test_merge_top_k():
    # src/collector/sort_key_top_collector.rs:168-192 (Desc = order_by_score ordering)
    rng = np.random.default_rng(3)
    vals = [(float(v), 0, v) for v in rng.permutation(10)]
    assert O.merge_top_k(vals, 0, 0) == [] and O.merge_top_k(vals, 3, 3) == []
    assert [(s, d) for s, _, d in O.merge_top_k(vals, 0, 2)] == [(9.0, 9), (8.0, 8)]
    assert [(s, d) for s, _, d in O.merge_top_k(vals, 2, 4)] == [(7.0, 7), (6.0, 6)]
    assert len(O.merge_top_k(vals, 0, 11)) == 10


# ---- positions codec (SURVEY.md §8f N3 groundwork): src/positions/mod.rs tests ------------------------------------
def _positions(vals, chunk=0):
    data = O.positions_serialize(vals, chunk)
    return data, O.PositionReader(data)


def test_positions_sizes_pinned_by_the_reference():
    # positions/mod.rs:82-96,131-134 (1000 deltas 0..999 -> 1224 bytes), :190-194 (512 -> 533),
    # :167-171 (2_000_000 -> 5_003_499), :204-209 (2_000_000 x 9 -> 1_015_627), :100-108 (empty term is readable)
    assert len(O.positions_serialize(np.arange(1000))) == 1224
    assert len(O.positions_serialize(np.arange(512))) == 533
    assert len(O.positions_serialize(np.arange(2_000_000))) == 5_003_499
    assert len(O.positions_serialize(np.full(2_000_000, 9))) == 1_015_627
    empty = O.positions_serialize([])
    assert bytes(empty) == b"\x80"  # VInt(0) bit-packed blocks, nothing else
    O.PositionReader(empty)
    # postings/mod.rs:61-81 test_position_write: 120 docs x deltas [1, 2, 3, 2] -> a 207-byte `.pos` FILE = this term's
    # 196 bytes (VInt(3) + 3 widths + 3 x 32-byte blocks of 2-bit deltas + 96 VInt bytes) + the 11-byte composite footer
    pw = bytes(O.positions_serialize(np.tile([1, 2, 3, 2], 120)))
    assert len(pw) == 207 - 11
    # the field split the size implies (positions/mod.rs:22-28): 480 deltas = 3 full blocks + 96 in the VInt tail
    assert pw[0] == 0x80 | 3                      # VInt(number of bit-packed blocks), stop bit on the last byte
    assert pw[1:4] == bytes([2, 2, 2])            # one bit-width byte per block: deltas <= 3 need 2 bits
    assert len(pw[4:4 + 3 * 32]) == 96            # 3 blocks x 16 * 2 bytes (BitPacker4x: 16 bytes per bit of width)
    assert pw[100:] == bytes([0x80 | v for v in [1, 2, 3, 2]] * 24)  # 96 one-byte VInts
    # what BitPacker4x's published layout says those 2-bit blocks hold: value j in bit stream j & 3 at bit 2 * (j >> 2);
    # the deltas repeat with period 4, so stream c holds the constant (1, 2, 3, 2)[c] in all its 16 fields
    words = np.frombuffer(pw[4:100], dtype="<u4").reshape(3, 2, 4)
    for c, v in enumerate([1, 2, 3, 2]):
        assert (words[:, :, c] == sum(v << (2 * i) for i in range(16))).all()


def test_positions_read_offsets_and_rereads():
    data, r = _positions(np.arange(1000))
    for n in (1, 10, 127, 128, 130, 312):  # mod.rs:82-96
        assert (r.read(0, n) == np.arange(n)).all()
    for offset in (1, 10, 127, 128, 130, 312):  # mod.rs:131-146
        for n in (1, 10, 130, 500):
            assert (r.read(offset, n) == np.arange(offset, offset + n)).all()
    _, r = _positions(np.arange(1000))
    c = 0
    for step in range(100):  # mod.rs:149-166 read twice, then skip
        a = r.read(7 * step, 7)
        b = r.read(7 * step, 7)
        assert (a == b).all() and (a == np.arange(c, c + 7)).all()
        c += 7
    _, r = _positions(np.arange(512))
    assert r.read(230, 1)[0] == 230 and r.read(9, 1)[0] == 9  # mod.rs:190-201 going back resets the reader
    _, r = _positions(np.arange(2_000_000))
    for _ in range(2):  # mod.rs:168-183 anchor != block
        assert (r.read(128, 256) == np.arange(128, 384)).all()
    data, _ = _positions(np.arange(2_000_000))
    for offset in (10, 128 * 1024, 128 * 1024 - 1, 128 * 1024 + 7, 128 * 10 * 1024 + 10):  # mod.rs:213-235
        assert O.PositionReader(data).read(offset, 1)[0] == offset


def test_positions_multiple_writes_and_random_round_trips():
    data = O.positions_serialize([1, 12, 4, 17, 443], chunk=2)  # mod.rs:110-129: several write_positions_delta calls
    assert (O.PositionReader(data).read(0, 5) == [1, 12, 4, 17, 443]).all()
    assert bytes(data) == bytes(O.positions_serialize([1, 12, 4, 17, 443]))
    rng = np.random.default_rng(21)
    for n in (0, 1, 70, 127, 128, 129, 200, 255, 256, 257, 270):  # mod.rs:57-66 proptest sizes
        vals = rng.choice([1, 2, 4, 8, 16], size=n).astype(np.uint32)
        data = O.positions_serialize(vals, chunk=int(rng.integers(1, 50)))
        r = O.PositionReader(data)
        for i in range(n):
            assert r.read(i, 1)[0] == vals[i]
        if n:
            assert (O.PositionReader(data).read(0, n) == vals).all()


def test_positions_of_the_compat_fixture():
    # tests/compat_tests_data/index_v{6,7}/*.pos: the label field's only term has one position, 0:
    # VInt(0 blocks) + VInt(0) = 80 80 (the TermInfo in the `.term` file says positions_range 0..2)
    for ver in ("index_v6", "index_v7"):
        pos = bytes.fromhex(GOLDEN["compat"][ver]["pos"])
        assert pos[:2] == bytes(O.positions_serialize([0])) == b"\x80\x80"
        assert O.PositionReader(np.frombuffer(pos[:2], dtype=np.uint8)).read(0, 1)[0] == 0
