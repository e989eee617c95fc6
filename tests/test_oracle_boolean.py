"""N4: the oracle's restatement of BooleanWeight::complex_scorer for term leaves with mixed Occurs
(oracle/oracle_api.cpp bool_for_each), pinned on the reference's own tests:
  src/query/boolean_query/mod.rs:109-170   test_boolean_query (doc sets of Must / Should / MustNot mixes)
  src/query/boolean_query/mod.rs:172-219   test_boolean_query_two_excluded (exclusion does not change the score)
  src/query/boolean_query/boolean_query.rs:287-340  test_minimum_required (minimum_number_should_match doc sets)
and differential-tested against set arithmetic and against the AND / OR paths it generalises."""
import numpy as np

from oracle import tq_oracle as O
from tantivy_b200._abi import (TQ_OCCUR_MUST, TQ_OCCUR_MUST_NOT, TQ_OCCUR_SHOULD, TQ_OP_AND, TQ_OP_BOOL, TQ_OP_OR, TQ_RECORD_FREQS, QueryBatch)
from tests.helpers import OracleSegment, hits, make_query
from tests.test_phrase_goldens import build

M, S, N = TQ_OCCUR_MUST, TQ_OCCUR_SHOULD, TQ_OCCUR_MUST_NOT


def text_segment(texts, segment_ord=0):
    postings, lengths = build(texts)
    vocab = sorted(postings)
    lists = [(np.array([d for d, _ in postings[t]], dtype=np.uint32), np.array([len(p) for _, p in postings[t]], dtype=np.uint32)) for t in vocab]
    return OracleSegment(lists, np.array(lengths), record_option=TQ_RECORD_FREQS, segment_ord=segment_ord), vocab


def bool_query(segs, clauses, k, msm=0):
    """clauses: list of (occur, term ordinal, group or None); absent terms (ordinal None) become clauses without postings."""
    terms = [t if t is not None else 0 for _, t, _ in clauses]
    q = make_query(TQ_OP_OR, segs, terms, k)
    q["term_segs"] = [ts for ts in q["term_segs"] if clauses[ts[0]][1] is not None]
    q["op"] = TQ_OP_BOOL
    q["term_occur"] = [o for o, _, _ in clauses]
    if any(g is not None for _, _, g in clauses):
        q["term_group"] = [g if g is not None else 200 + i for i, (_, _, g) in enumerate(clauses)]
    q["min_should_match"] = msm
    return q


def docs_of(ix, q, i=0):
    return sorted(d for _, _, d in hits(ix.search_batch(QueryBatch([q]), mode=0), i))


def test_reference_doc_sets():
    seg, vocab = text_segment(["a b c", "a c", "b c", "a b c d", "d"])
    ix = O.OracleIndex()
    seg.register(ix)
    t = {w: vocab.index(w) for w in vocab}
    q = lambda clauses, msm=0: docs_of(ix, bool_query([seg], clauses, 10, msm))  # noqa: E731
    assert q([(M, t["a"], None)]) == [0, 1, 3]
    assert q([(S, t["a"], None)]) == [0, 1, 3]
    assert q([(S, t["a"], None), (S, t["b"], None)]) == [0, 1, 2, 3]
    assert q([(M, t["a"], None), (S, t["b"], None)]) == [0, 1, 3]
    assert q([(M, t["a"], None), (S, t["b"], None), (N, t["d"], None)]) == [0, 1]
    assert q([(N, t["d"], None)]) == []
    # +a +(b c)  (mod.rs:76-81): an Intersection of a term and a union
    assert q([(M, t["a"], 0), (M, t["b"], 1), (M, t["c"], 1)]) == [0, 1, 3]
    # two_excluded: the score of doc 4 for +d is not changed by -a -b
    r1 = hits(ix.search_batch(QueryBatch([bool_query([seg], [(M, t["d"], None)], 3)]), mode=0))
    r2 = hits(ix.search_batch(QueryBatch([bool_query([seg], [(M, t["d"], None), (N, t["a"], None), (N, t["b"], None)], 3)]), mode=0))
    assert [d for _, _, d in r1] == [4, 3] and [d for _, _, d in r2] == [4] and r1[0][0] == r2[0][0]


def test_reference_minimum_required():
    seg, vocab = text_segment(["a b c", "a c e", "d f g", "z z z", "c i b"])
    ix = O.OracleIndex()
    seg.register(ix)
    t = {w: vocab.index(w) for w in vocab}
    q = lambda words, mr: docs_of(ix, bool_query([seg], [(S, t.get(w), None) for w in words], 10, mr))  # noqa: E731
    assert q(["a", "c", "z", "i"], 2) == [0, 1, 4]
    assert q(["a", "b", "c", "e"], 3) == [0, 1]
    assert q(["a", "b"], 3) == []
    assert q(["a", "b"], 2) == [0]      # as many as there are clauses: they act as MUST clauses
    assert q(["a", "zzzz"], 1) == [0, 1]  # an absent term is an EmptyScorer: removed before the count (boolean_weight.rs:247-263)
    assert q(["a", "zzzz"], 2) == []


def _random_segments(rng, n):
    segs = []
    for so in range(n):
        max_doc = int(rng.integers(500, 5000))
        lengths = np.clip(np.round(np.exp(rng.normal(np.log(40), 0.7, size=max_doc))), 1, 4096).astype(np.uint32)
        lists = []
        for p in (0.5, 0.2, 0.1, 0.05, 0.02, 0.3):
            docs = np.nonzero(rng.random(max_doc) < p)[0].astype(np.uint32)
            lists.append((docs, np.minimum(rng.geometric(0.6, size=len(docs)), 10).astype(np.uint32)))
        segs.append(OracleSegment(lists, lengths, segment_ord=so))
    return segs


def test_boolean_generalises_and_or_and_set_arithmetic():
    rng = np.random.default_rng(77)
    segs = _random_segments(rng, 2)
    ix = O.OracleIndex()
    for s in segs:
        s.register(ix)
    big = 100_000
    # pure shapes: bit-identical rows to the AND / OR paths
    for terms in ([0, 1], [2, 0, 4], [5, 3]):
        a = hits(ix.search_batch(QueryBatch([make_query(TQ_OP_AND, segs, terms, 50)]), mode=0))
        b = hits(ix.search_batch(QueryBatch([bool_query(segs, [(M, t, None) for t in terms], 50)]), mode=0))
        assert a == b
        a = hits(ix.search_batch(QueryBatch([make_query(TQ_OP_OR, segs, terms, 50)]), mode=0))
        b = hits(ix.search_batch(QueryBatch([bool_query(segs, [(S, t, None) for t in terms], 50)]), mode=0))
        assert a == b
    # mixed shapes of benches/and_or_queries.rs:142-155 and more: doc sets by set arithmetic, per segment
    shapes = [
        [(M, 2, 0), (M, 1, 1), (M, 3, 1)],                      # +c +(b OR d)
        [(M, 0, 0), (M, 2, 1), (M, 4, 1)],                      # +e +(c OR a)
        [(M, 2, 0), (M, 1, 0), (M, 3, 1), (M, 0, 1)],           # +(c OR b) +(d OR e)
        [(M, 0, None), (S, 1, None), (S, 4, None)],             # +a b e
        [(S, 1, None), (S, 2, None), (N, 0, None)],             # b c -a
        [(M, 5, None), (N, 0, None), (N, 3, None)],             # +f -a -d
        [(M, 1, 7), (M, 2, 7), (S, 3, None), (N, 4, None)],     # +(b OR c) d -e
    ]
    for clauses in shapes:
        got = hits(ix.search_batch(QueryBatch([bool_query(segs, clauses, big)]), mode=0))
        want = set()
        for s in segs:
            docsets = {}
            for _, t, _ in clauses:
                df, a, e = s.terms[t]
                docsets[t] = set(ix.decode_postings(s.term_seg(t))[0].tolist()) if df else set()
            groups = {}
            for i, (o, t, g) in enumerate(clauses):
                if o == M:
                    groups.setdefault(g if g is not None else 1000 + i, set()).update(docsets[t])
            shoulds = [docsets[t] for o, t, _ in clauses if o == S]
            nots = [docsets[t] for o, t, _ in clauses if o == N]
            cand = set.intersection(*groups.values()) if groups else set.union(*shoulds)
            for x in nots:
                cand -= x
            want |= {(s.segment_ord, d) for d in cand}
        assert {(g, d) for _, g, d in got} == want, clauses
        assert ix.count_batch(QueryBatch([bool_query(segs, clauses, 1)]))[0] == len(want)
    # minimum_should_match with a MUST: docs of a that also hold at least two of b, c, d
    clauses = [(M, 0, None), (S, 1, None), (S, 2, None), (S, 3, None)]
    got = hits(ix.search_batch(QueryBatch([bool_query(segs, clauses, big, msm=2)]), mode=0))
    want = set()
    for s in segs:
        ds = [set(ix.decode_postings(s.term_seg(t))[0].tolist()) for t in (0, 1, 2, 3)]
        want |= {(s.segment_ord, d) for d in ds[0] if sum(d in x for x in ds[1:]) >= 2}
    assert {(g, d) for _, g, d in got} == want
