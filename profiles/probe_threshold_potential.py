"""Upper bound for a threshold pre-pass: run the default union batch, then run it again with each query's final k-th
score (one ulp below) handed in as tq_query.threshold, and compare kernel times and the window routes of k_or_strip
(or_windows[1] exhaustive, [2] hot = non-essential clauses applied, [3] cold = essential clauses only)."""
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
for it in range(2):
    res = ctx.search_batch(qb)
    st = ctx.stats()
    print("plain     or_ms", round(st["or_ms"], 1), "windows", st["or_windows"][:5])
w0 = np.array(st["or_windows"], dtype=np.int64)
scores, counts = res[0], res[3]
thr = np.array([np.nextafter(scores[i, counts[i] - 1], np.float32(-np.inf), dtype=np.float32) if counts[i] else np.float32(0) for i in range(nq)])
frac = float(os.environ.get("THR_FRAC", "1.0"))
qb.q["flags"] = 1
qb.q["threshold"] = thr * np.float32(frac)
for it in range(2):
    res2 = ctx.search_batch(qb)
    st = ctx.stats()
    w1 = np.array(st["or_windows"], dtype=np.int64)
    print("threshold or_ms", round(st["or_ms"], 1), "windows of this run", list((w1 - w0)[:5]))
    w0 = w1
assert (res2[3] == res[3]).all() and (res2[2] == res[2]).all() and (res2[0] == res[0]).all(), "results changed"
print("results identical")
