"""N3 groundwork: the reference's exact-phrase semantics, pinned on its own tests before a kernel
exists.  What a device kernel will have to reproduce, restated here from
  src/query/phrase_query/phrase_scorer.rs:349-398,431-497  every term's positions are shifted by (max_offset - its
                                                           offset); phrase_count = size of the intersection of the
                                                           shifted position sets (slop = 0)
  src/query/bm25.rs:95-129                                 Bm25Weight::for_terms: weight = (sum of the terms' idf) * (1+K1)
  src/query/phrase_query/phrase_scorer.rs:576-589          score = weight * tf_factor(fieldnorm_id, phrase_count)
and checked against src/query/phrase_query/mod.rs:41-73,76-91,163-169 (doc sets and scores).  The position streams the
kernel will read are round-tripped through the oracle's position codec (oracle/positions.hpp)."""
import numpy as np

from oracle import tq_oracle as O
from tests.helpers import f32, fieldnorm_ids


def tokenize(text):
    return text.lower().split()


def build(texts):
    """term -> [(doc, positions)] and doc lengths, as the reference's indexer produces them."""
    postings, lengths = {}, []
    for doc, text in enumerate(texts):
        toks = tokenize(text)
        lengths.append(len(toks))
        for pos, t in enumerate(toks):
            pl = postings.setdefault(t, [])
            if pl and pl[-1][0] == doc:
                pl[-1][1].append(pos)
            else:
                pl.append((doc, [pos]))
    return postings, lengths


def positions_via_codec(plist):
    """The term's `.pos` bytes (delta per doc: first position, then gaps) written and read back by the oracle codec."""
    deltas = []
    for _, ps in plist:
        deltas += [ps[0]] + [b - a for a, b in zip(ps, ps[1:])]
    reader = O.PositionReader(O.positions_serialize(deltas))
    out, offset = [], 0
    for doc, ps in plist:
        d = reader.read(offset, len(ps))  # segment_postings.rs positions_with_offset: tf deltas at the running offset
        offset += len(ps)
        out.append((doc, list(np.cumsum(d))))
    return out


def phrase_search(texts, phrase):
    postings, lengths = build(texts)
    n_docs = len(texts)
    avg = f32(f32(sum(lengths)) / f32(n_docs))
    if any(t not in postings for t in phrase):
        return []
    lists = [dict(positions_via_codec(postings[t])) for t in phrase]
    max_offset = len(phrase) - 1
    idf_sum = f32(0)
    for t in phrase:
        idf_sum = f32(idf_sum + O.bm25_idf(len(postings[t]), n_docs))
    weight = f32(idf_sum * f32(2.2))
    cache = O.bm25_tf_cache(avg)
    ids = fieldnorm_ids(lengths)
    hits = []
    for doc in sorted(set.intersection(*[set(l) for l in lists])):
        shifted = [set(p + (max_offset - off) for p in l[doc]) for off, l in enumerate(lists)]
        count = len(set.intersection(*shifted))
        if count:
            tf = f32(count)
            hits.append((doc, float(f32(weight * f32(tf / f32(tf + cache[ids[doc]]))))))
    return hits


def test_phrase_query_doc_sets():  # phrase_query/mod.rs:41-73
    texts = ["b b b d c g c", "a b b d c g c", "a b a b c", "c a b a d ga a", "a b c"]
    docs = lambda phrase: [d for d, _ in phrase_search(texts, phrase)]  # noqa: E731
    assert docs(["a", "b"]) == [1, 2, 3, 4]
    assert docs(["a", "b", "c"]) == [2, 4]
    assert docs(["b", "b"]) == [0, 1]
    assert docs(["g", "ewrwer"]) == []
    assert docs(["g", "a"]) == []


def test_phrase_query_simple():  # phrase_query/mod.rs:76-91
    assert [d for d, _ in phrase_search(["a b b d c g c", "a b a b c"], ["a", "b"])] == [0, 1]


def test_phrase_score():  # phrase_query/mod.rs:163-169
    hits = phrase_search(["a b c", "a b c a b"], ["a", "b"])
    assert [d for d, _ in hits] == [0, 1]
    assert abs(hits[0][1] - 0.40618482) <= 1e-6 and abs(hits[1][1] - 0.46844664) <= 1e-6
