"""The C++ host mirror of the reference's search API (tantivy_b200/host/tantivy_host.hpp).

The mirror's own tests are C++ (tests/cpp/host_mirror_tests.cpp, in the shape of the reference's tests); this module
  * runs their host-only part on the CPU box (tokenizer, fieldnorms, statistics, composite-file / footer framing of
    the reference's compat fixture, error kinds, "no CUDA device -> raises"),
  * feeds the segments the mirror's IndexWriter produced to the ORACLE and checks the reference's golden scores
    there (term_query/mod.rs:21-129, boolean_query/mod.rs:221-291, top_score_collector.rs:838-921), which pins the
    indexing chain (tokenizer -> postings -> serializer) without a GPU,
  * and, on the GPU box, runs the C++ search tests through the C ABI and cross-checks the same segments
    GPU-vs-oracle from Python."""
import json
import os
import subprocess

import numpy as np
import pytest

from oracle import tq_oracle as O
from tantivy_b200._abi import TQ_OP_AND, TQ_OP_OR, TQ_OP_TERM, QueryBatch
from tests.helpers import GOLDEN, f32, hits

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "tantivy_b200", "_lib", "host_mirror_tests")


def _run(args, **kw):
    return subprocess.run([BIN] + args, capture_output=True, text=True, timeout=300, **kw)


@pytest.fixture(scope="module")
def workdir(tmp_path_factory):
    d = tmp_path_factory.mktemp("host_mirror")
    for version, fx in GOLDEN["compat"].items():  # the reference's index_v6 / index_v7 directories, rebuilt from the fixture
        sub = d / version
        sub.mkdir()
        (sub / "meta.json").write_text(json.dumps(fx["meta"], indent=2))
        stem = fx["meta"]["segments"][0]["segment_id"].replace("-", "")
        for ext in ("idx", "fieldnorm", "pos", "term"):
            (sub / f"{stem}.{ext}").write_bytes(bytes.fromhex(fx[ext]))
    r = _run(["--dump", str(d)])
    assert r.returncode == 0, r.stdout + r.stderr
    return d


class DumpedField:
    """One field of one dumped segment, shaped like tests.helpers.OracleSegment."""

    def __init__(self, d, seg, f, spec, segment_ord):
        self.segment_ord, self.field = segment_ord, f
        self.max_doc = seg["max_doc"]
        self.record_option = spec["record"]
        self.body = np.frombuffer((d / spec["idx"]).read_bytes(), dtype=np.uint8)
        self.fn_ids = np.frombuffer((d / spec["fieldnorm"]).read_bytes(), dtype=np.uint8) if spec["fieldnorm"] else None
        self.alive = np.frombuffer((d / seg["alive"]).read_bytes(), dtype=np.uint8) if seg["alive"] else None
        self.total_num_tokens = int(np.frombuffer(self.body[:8].tobytes(), dtype="<u8")[0])
        self.terms = {t: tuple(v[:3]) for t, v in spec["terms"].items()}
        self.positions_range = {t: (v[3], v[4]) for t, v in spec["terms"].items()}
        self.pos = np.frombuffer((d / spec["idx"].replace(".idx", ".pos")).read_bytes(), dtype=np.uint8)

    def register(self, index):
        index.segment_register(self.segment_ord, self.field, self.max_doc, self.record_option, self.body, self.fn_ids, self.alive)


def load(d, name):
    m = json.loads((d / "manifest.json").read_text())[name]
    return {fname: [DumpedField(d, seg, f, seg["fields"][f], so) for so, seg in enumerate(m["segments"])]
            for f, fname in enumerate(m["fields"])}


def query(op, segs, terms, k, boost=1.0, flags=None):
    """Bm25Weight::for_terms over the searcher's statistics (bm25.rs:27-50,95-118)."""
    n_docs = sum(s.max_doc for s in segs)
    avg = f32(f32(sum(s.total_num_tokens for s in segs)) / f32(n_docs))
    weights, term_segs = [], []
    for clause, t in enumerate(terms):
        df = sum(s.terms.get(t, (0, 0, 0))[0] for s in segs)
        weights.append(O.bm25_weight(df, n_docs, boost))
        for s in segs:
            if t in s.terms:
                term_segs.append((clause, s.segment_ord, s.field) + s.terms[t])
    q = dict(op=op, k=k, weights=weights, avg_fieldnorm=[avg] * len(terms), term_segs=term_segs)
    if flags:
        q["term_flags"] = flags
    return q


def oracle_search(segs, q, mode=0):
    ix = O.OracleIndex()
    for s in segs:
        s.register(ix)
    return hits(ix.search_batch(QueryBatch([q]), mode=mode))


def near(a, b):
    return abs(a - b) <= 1e-6 * max(abs(a), abs(b))


def test_binary_is_built():
    assert os.path.exists(BIN), "run __graft_entry__.build() (make -C tantivy_b200/csrc)"


def test_host_side_checks(workdir):
    r = _run(["--cpu", str(workdir)])
    assert r.returncode == 0, r.stdout + r.stderr
    assert "ok   host_search_without_device_raises" in r.stdout or "ok   host_tokenizer_and_statistics" in r.stdout
    assert "ok   host_compat_framing" in r.stdout and "ok   host_term_info_store" in r.stdout
    assert "ok   host_fst_term_dictionary" in r.stdout  # N2, FST kind: the fixtures' dictionaries + self-consistency of the restated layout
    assert "ok   host_sstable_term_dictionary" in r.stdout  # N2, SSTable kind: sstable/'s golden bytes, `.term` of that kind opened
    assert "ok   host_alive_bitset_file" in r.stdout  # `.del`: BitSet::serialize + footer, a segment opened with deletes


def test_mirror_segments_reproduce_reference_scores_in_the_oracle(workdir):
    # term_query/mod.rs:21-44
    segs = load(workdir, "one_doc_string")["text"]
    h = oracle_search(segs, query(TQ_OP_TERM, segs, ["a"], 1))
    assert [(g, d) for _, g, d in h] == [(0, 0)] and near(h[0][0], 0.28768212)
    # term_query/mod.rs:80-129
    segs = load(workdir, "term_weight")["left"]
    h = oracle_search(segs, query(TQ_OP_TERM, segs, ["left2"], 2))
    assert len(h) == 1 and near(h[0][0], 0.77802235)
    h = oracle_search(segs, query(TQ_OP_TERM, segs, ["left1"], 2))
    assert len(h) == 2 and near(h[0][0], 0.27101856) and near(h[1][0], 0.13736556)
    h = oracle_search(segs, query(TQ_OP_OR, segs, ["left2", "left1"], 2))
    assert len(h) == 2 and near(h[0][0], 0.9153879) and near(h[1][0], 0.27101856)
    # boolean_query/mod.rs:221-259
    segs = load(workdir, "boolean_weight")["text"]
    h = oracle_search(segs, query(TQ_OP_OR, segs, ["a", "b"], 3))
    assert (h[0][1], h[0][2]) == (0, 0) and near(h[0][0], 0.84163445)
    h = oracle_search(segs, query(TQ_OP_OR, segs, ["a", "b"], 3, boost=2.0))
    assert near(h[0][0], 1.6832689)
    # boolean_query/mod.rs:262-291 (IndexRecordOption::Basic requested on a TEXT field)
    segs = load(workdir, "boolean_aux")["text"]
    h = oracle_search(segs, query(TQ_OP_AND, segs, ["a", "b"], 10, flags=[1, 1]))
    assert [(d, ) for _, _, d in h] == [(0, ), (3, )] and near(h[0][0], 0.977973) and near(h[1][0], 0.84699446)
    # top_score_collector.rs:838-857
    segs = load(workdir, "droopy")["text"]
    h = oracle_search(segs, query(TQ_OP_OR, segs, ["droopy", "tax"], 4))
    assert [d for _, _, d in h] == [1, 2, 0]
    assert near(h[0][0], 0.81221175) and near(h[1][0], 0.5376842) and near(h[2][0], 0.48527452)
    for mode in (0, 1):  # the reference-faithful pruned path agrees
        assert oracle_search(segs, query(TQ_OP_OR, segs, ["droopy", "tax"], 2), mode)[0][2] == 1


def test_mirror_multi_segment_index_is_well_formed(workdir):
    segs = load(workdir, "multi_segment")["body"]
    assert [s.max_doc for s in segs] == [700, 811, 922]
    assert segs[0].alive is not None and segs[1].alive is not None and segs[2].alive is None
    ix = O.OracleIndex()
    for s in segs:
        s.register(ix)
    for s in segs:  # every list decodes to doc_freq strictly ascending docs below max_doc
        for t, (df, st, en) in list(s.terms.items())[:12]:
            docs, tfs = ix.decode_postings((0, s.segment_ord, s.field, df, st, en))
            assert len(docs) == df and (np.diff(docs.astype(np.int64)) > 0).all() and docs[-1] < s.max_doc and (tfs >= 1).all()
    rare = query(TQ_OP_TERM, segs, ["rare"], 1000)
    h = oracle_search(segs, rare)
    assert h and all(g == 2 for _, g, _ in h)  # deleted in segments 0 and 1
    a = oracle_search(segs, query(TQ_OP_OR, segs, ["w0", "w3", "w11"], 300), 0)
    b = oracle_search(segs, query(TQ_OP_OR, segs, ["w0", "w3", "w11"], 300), 1)
    assert len(a) == 300 and [(g, d) for _, g, d in a[:50]] == [(g, d) for _, g, d in b[:50]]


def test_mirror_writes_the_reference_position_format(workdir):
    """`.pos` of the mirror's IndexWriter (tq::encode_positions): byte-identical to the oracle's PositionSerializer on a
    text whose positions are known, and self-consistent (tf positions per posting, increasing, below the doc length)
    on the pseudo-random three-segment index."""
    segs = load(workdir, "droopy")["text"]
    texts = ["hello happy tax payer", "droopy says hello happy tax payer", "i like droopy"]
    seg = segs[0]
    for term, (a, b) in seg.positions_range.items():
        deltas = []
        for text in texts:
            ps = [i for i, t in enumerate(text.split()) if t == term]
            if ps:
                deltas += [ps[0]] + [y - x for x, y in zip(ps, ps[1:])]
        assert bytes(seg.pos[a:b]) == bytes(O.positions_serialize(deltas)), term
    ix = O.OracleIndex()
    segs = load(workdir, "multi_segment")["body"]
    for s in segs:
        s.register(ix)
    for s in segs:
        for term in ("w0", "w3", "rare", "w16"):
            if term not in s.terms:
                continue
            df, st, en = s.terms[term]
            docs, tfs = ix.decode_postings((0, s.segment_ord, s.field, df, st, en))
            a, b = s.positions_range[term]
            reader = O.PositionReader(s.pos[a:b])
            deltas = reader.read(0, int(tfs.sum()))
            assert bytes(O.positions_serialize(deltas)) == bytes(s.pos[a:b])  # nothing but these deltas in the range
            off = 0
            for d, tf in zip(docs[:200], tfs[:200]):
                ps = np.cumsum(deltas[off:off + tf])
                off += int(tf)
                assert (np.diff(ps) > 0).all() and ps[-1] < 64  # docs of this index have at most 31 + 1 tokens


def _phrase(workdir, name, words, slop=0):
    """PhraseQuery::new(terms) with set_slop(slop) on a dumped index, through the oracle's PhraseScorer
    (phrase_query/mod.rs:203-218 test_query): [(doc, score)] in doc order."""
    segs = load(workdir, name)["text"]
    ix = O.OracleIndex()
    for s in segs:
        s.register(ix)
        ix.register_positions(s.segment_ord, s.field, s.pos)
    n_docs = sum(s.max_doc for s in segs)
    avg = f32(f32(sum(s.total_num_tokens for s in segs)) / f32(n_docs))
    idf_sum = f32(0)
    terms = []
    for off, wd in enumerate(words):
        df = sum(s.terms.get(wd, (0, 0, 0))[0] for s in segs)
        if df == 0:
            return []
        idf_sum = f32(idf_sum + O.bm25_idf(df, n_docs))  # Bm25Weight::for_terms (bm25.rs:95-129)
        for s in segs:
            if wd in s.terms:
                terms.append((off, s.segment_ord, s.field) + s.terms[wd] + s.positions_range[wd])
    return [(d, sc) for _, d, sc, _ in ix.phrase_search(terms, f32(idf_sum * f32(2.2)), avg, slop=slop)]


def test_oracle_phrase_scorer_reproduces_the_reference_phrase_tests(workdir):
    docs = lambda name, words, slop=0: [d for d, _ in _phrase(workdir, name, words, slop)]  # noqa: E731
    # phrase_query/mod.rs:41-73
    assert docs("phrase_query", ["a", "b"]) == [1, 2, 3, 4]
    assert docs("phrase_query", ["a", "b", "c"]) == [2, 4]
    assert docs("phrase_query", ["b", "b"]) == [0, 1]
    assert docs("phrase_query", ["g", "ewrwer"]) == [] and docs("phrase_query", ["g", "a"]) == []
    assert docs("phrase_simple", ["a", "b"]) == [0, 1]  # :76-91
    # :163-169
    h = _phrase(workdir, "phrase_score", ["a", "b"])
    assert [d for d, _ in h] == [0, 1] and near(h[0][1], 0.40618482) and near(h[1][1], 0.46844664)
    # slop: :182-201, :220-226
    assert len(_phrase(workdir, "phrase_slop_bug", ["captain", "wendy"], 1)) == 1
    assert len(_phrase(workdir, "phrase_slop_bug_2a", ["a", "b", "c"], 2)) == 1
    assert len(_phrase(workdir, "phrase_slop_bug_2b", ["a", "b", "c"], 2)) == 1
    assert len(_phrase(workdir, "phrase_slop_repeating", ["wendy", "subject", "captain"], 1)) == 1
    # :228-235
    h = _phrase(workdir, "phrase_slop_size", ["a", "c"], 3)
    assert len(h) == 2 and near(h[0][1], 0.29086056) and near(h[1][1], 0.26706287)
    # :238-256
    assert len(_phrase(workdir, "phrase_slop_1", ["a", "b", "c"], 1)) == 1
    assert len(_phrase(workdir, "phrase_slop_2", ["a", "b", "c"], 1)) == 0
    assert len(_phrase(workdir, "phrase_slop_3", ["b", "a"], 1)) == 0
    assert len(_phrase(workdir, "phrase_slop_3", ["b", "a"], 2)) == 1
    # :259-274
    h = _phrase(workdir, "phrase_slop_ordering", ["a", "b", "c"], 3)
    assert near(h[0][1], 0.23091172) and near(h[1][1], 0.27310878) and near(h[3][1], 0.25024384)


@pytest.mark.gpu
def test_cpp_search_tests_on_the_device(workdir):
    r = _run(["--compat", str(workdir)])
    print(r.stdout)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "ok   compat_index_search" in r.stdout and "0 test(s) failed" in r.stdout


@pytest.mark.gpu
def test_mirror_segments_gpu_vs_oracle(workdir):
    import tantivy_b200 as T
    segs = load(workdir, "multi_segment")["body"]
    for s in segs:
        s.segment_ord += 7000  # private ordinals inside this process
    ctx, ix = T.Context(0), O.OracleIndex()
    for s in segs:
        s.register(ctx)
        s.register(ix)
    vocab = sorted(segs[0].terms)
    queries = [query(TQ_OP_TERM, segs, [t], 40) for t in vocab[:8]]
    queries += [query(TQ_OP_OR, segs, ts, k) for ts in (vocab[:3], vocab[3:9], ["rare", "w0"]) for k in (10, 200)]
    queries += [query(TQ_OP_AND, segs, ts, 50) for ts in (["w0", "w1"], ["rare", "w2", "w3"])]
    queries += [query(TQ_OP_OR, segs, ["w0", "w1"], 30, flags=[1, 0])]
    qb = QueryBatch(queries)
    g, c = ctx.search_batch(qb), ix.search_batch(qb, mode=0)
    for i in range(qb.nq):
        assert hits(g, i) == hits(c, i), f"query {i}"
    ctx.close()
