"""How much of the default union workload could MaxScore skip?  For every query: the final threshold (100th score),
the clauses' score upper bounds (weight x largest tf factor), the non-essential prefix (clauses, by ascending bound,
whose bounds sum below the threshold) and the share of the query's postings that sit in non-essential lists."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import bench as B
import tantivy_b200 as T

wl = dict(B.WORKLOADS["or5_top100_100M_8seg"])
nq = 256
dens, batches = B.build_query_plan(wl, nq, 1, 0x7A6E7469)
shard = B.make_shard(wl, dens, 0, 1, 0x7A6E7469)
ctx = T.Context(0)
shard.register(ctx)
qb = B.marshal(shard, batches[0])
res = ctx.search_batch(qb)
scores, counts = res[0], res[3]
n_docs = shard.total_docs
tot_post = skip_post = 0
fr = []
for i, (op, terms, k) in enumerate(batches[0]):
    theta = float(scores[i, counts[i] - 1]) if counts[i] >= k else 0.0
    df = np.array([int(shard.df[t]) for t in terms], dtype=np.float64)
    w = np.array([T.bm25_weight(int(shard.df[t]), n_docs, 1.0) for t in terms], dtype=np.float64)
    bound = w * 0.97  # tf <= 10, shortest docs: tf / (tf + norm_min) stays below ~0.97
    order = np.argsort(bound)
    acc, ne = 0.0, []
    for j in order:
        if acc + bound[j] < theta:
            acc += bound[j]
            ne.append(j)
        else:
            break
    sk = df[ne].sum() if ne else 0.0
    tot_post += df.sum()
    skip_post += sk
    fr.append(sk / df.sum())
fr = np.array(fr)
print("queries", nq, "postings/query", tot_post / nq / 1e6, "M; skippable share (posting weighted)", skip_post / tot_post)
print("per-query skippable share: mean", fr.mean(), "median", np.median(fr), "p10", np.percentile(fr, 10), "p90", np.percentile(fr, 90))
print("queries with >50% skippable:", (fr > 0.5).mean(), " with 0:", (fr == 0).mean())
