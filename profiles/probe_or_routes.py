import sys, json, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench as B
import tantivy_b200 as T
wl = dict(B.WORKLOADS["or5_top100_100M_8seg"])
dens, batches = B.build_query_plan(wl, 256, 2, 0x7A6E7469)
shard = B.make_shard(wl, dens, 0, 1, 0x7A6E7469)
ctx = T.Context(0); shard.register(ctx)
qbs = [B.marshal(shard, b) for b in batches]
for it in range(3):
    t0=time.time(); ctx.search_batch(qbs[it%2]); dt=time.time()-t0
    st = ctx.stats(); print("prune", os.environ.get("TQ_OR_PRUNE","1"), "iter", it, "ms", round(dt*1e3,1), "or_ms", round(st["or_ms"],1), "windows", st["or_windows"])
