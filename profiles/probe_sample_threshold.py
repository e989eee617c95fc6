"""What does a sampled threshold buy?  Thresholds = each query's k-th best score over a SAMPLE of the docs (the first
n of the 8 segments, searched exhaustively), then the full batch runs with them as tq_query.threshold."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import bench as B
import tantivy_b200 as T
from tantivy_b200._abi import QueryBatch
from tantivy_b200.lib import bm25_weight

wl = dict(B.WORKLOADS["or5_top100_100M_8seg"])
nq = 256
dens, batches = B.build_query_plan(wl, nq, 1, 0x7A6E7469)
shard = B.make_shard(wl, dens, 0, 1, 0x7A6E7469)
ctx = T.Context(0)
shard.register(ctx)


def marshal(queries, segments=None, thresholds=None):
    out = []
    for qi, (op, terms, k) in enumerate(queries):
        weights = [bm25_weight(int(shard.df[t]), shard.total_docs, 1.0) for t in terms]
        term_segs = []
        for clause, t in enumerate(terms):
            for i, g in enumerate(shard.global_ords):
                if segments is not None and g not in segments:
                    continue
                d, st, en = shard.ix.term_info[i][t]
                if d:
                    term_segs.append((clause, g, 0, d, st, en))
        q = dict(op=B.OPS[op], k=k, weights=weights, avg_fieldnorm=[shard.avg] * len(terms), term_segs=term_segs)
        if thresholds is not None:
            q["threshold"] = float(thresholds[qi])
        out.append(q)
    return QueryBatch(out)


full = marshal(batches[0])
for it in range(2):
    ref = ctx.search_batch(full)
    st = ctx.stats()
print("full, no threshold: or_ms", round(st["or_ms"], 1))
w0 = np.array(st["or_windows"], dtype=np.int64)
for n_sample in (1, 2):
    sample = marshal(batches[0], segments=set(range(n_sample)))
    for it in range(2):
        rs = ctx.search_batch(sample)
        st = ctx.stats()
    w0 = np.array(st["or_windows"], dtype=np.int64)
    print(f"sample = {n_sample}/8 of the docs: or_ms", round(st["or_ms"], 1))
    thr = np.array([np.nextafter(rs[0][i, rs[3][i] - 1], np.float32(-np.inf), dtype=np.float32) if rs[3][i] >= 100 else np.float32(-1)
                    for i in range(nq)])
    rest = marshal(batches[0], thresholds=thr)
    for it in range(2):
        r2 = ctx.search_batch(rest)
        st = ctx.stats()
        w1 = np.array(st["or_windows"], dtype=np.int64)
        print("   full run with the sampled thresholds: or_ms", round(st["or_ms"], 1), "windows [exh, hot, cold]", list((w1 - w0)[1:4]))
        w0 = w1
    assert (r2[2] == ref[2]).all() and (r2[0] == ref[0]).all()
print("results identical")
