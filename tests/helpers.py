"""Shared helpers of the test-suite (oracle-side segment construction, like the reference's
test-only constructors SegmentPostings::create_from_docs_and_tfs / TermScorer::create_for_test,
src/postings/segment_postings.rs:94, src/query/term_query/term_scorer.rs:33)."""
import json
import os

import numpy as np

from oracle import tq_oracle as O
from tantivy_b200._abi import TQ_OP_AND, TQ_OP_OR, TQ_OP_TERM, TQ_RECORD_FREQS, QueryBatch

GOLDEN = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_fixtures.json")))


def f32(x):
    return np.float32(x)


_FN_TABLE = np.array(GOLDEN["field_norms_table"], dtype=np.uint64)


def fieldnorm_ids(fieldnorms):
    """fieldnorm_to_id for an array: index of the last table entry <= fieldnorm (code.rs:7-11)."""
    f = np.asarray(fieldnorms, dtype=np.uint64)
    return (np.searchsorted(_FN_TABLE, f, side="right") - 1).astype(np.uint8)


class OracleSegment:
    """One segment, one field, built from explicit posting lists (docs, tfs) and doc lengths."""

    def __init__(self, posting_lists, fieldnorms, record_option=TQ_RECORD_FREQS, segment_ord=0, field=0,
                 writer_cls=None, alive=None, max_doc=None):
        if fieldnorms is None:  # field without fieldnorms: constant fieldnorm 1 (term_weight.rs:218)
            self.fieldnorms, self.fn_ids = None, None
            self.max_doc = max_doc
            self.total_num_tokens = max_doc
        else:
            self.fieldnorms = np.asarray(fieldnorms, dtype=np.uint32)
            self.max_doc = len(self.fieldnorms)
            self.fn_ids = fieldnorm_ids(self.fieldnorms)
            self.total_num_tokens = int(self.fieldnorms.astype(np.uint64).sum())
        self.record_option = record_option
        self.segment_ord, self.field = segment_ord, field
        writer_cls = writer_cls or O.FieldWriter
        w = writer_cls(record_option, self.total_num_tokens, self.fn_ids, self.max_doc)
        self.terms = []
        for docs, tfs in posting_lists:
            s, e = w.add_term(docs, tfs)
            self.terms.append((len(docs), s, e))
        self.body = w.body()
        self.alive = alive

    def term_seg(self, term, clause=0):
        df, s, e = self.terms[term]
        return (clause, self.segment_ord, self.field, df, s, e)

    def register(self, index):
        index.segment_register(self.segment_ord, self.field, self.max_doc, self.record_option, self.body, self.fn_ids, self.alive)


def searcher_stats(segments):
    """Bm25StatisticsProvider for Searcher (src/query/bm25.rs:27-50): N = sum max_doc,
    avg = total_num_tokens / N as f32."""
    n_docs = sum(s.max_doc for s in segments)
    tokens = sum(s.total_num_tokens for s in segments)
    return n_docs, f32(f32(tokens) / f32(n_docs))


def make_query(op, segments, terms, k, boost=1.0):
    """terms = list of term ordinals (same ordinal in every segment)."""
    n_docs, avg = searcher_stats(segments)
    weights, term_segs = [], []
    for clause, t in enumerate(terms):
        df = sum(s.terms[t][0] for s in segments)
        weights.append(O.bm25_weight(df, n_docs, boost))
        for s in segments:
            if s.terms[t][0] > 0:
                term_segs.append(s.term_seg(t, clause))
    return dict(op=op, k=k, weights=weights, avg_fieldnorm=[avg] * len(terms), term_segs=term_segs)


def hits(result, i=0):
    scores, segs, docs, counts = result
    n = int(counts[i])
    return [(float(scores[i, j]), int(segs[i, j]), int(docs[i, j])) for j in range(n)]
