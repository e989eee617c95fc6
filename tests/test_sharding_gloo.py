"""world_size-2 test of the multi-GPU host logic on CPU (gloo): segment assignment, global BM25
statistics by all_reduce, per-rank search, all_gather of the rows and merge_fruits across ranks.
The per-rank search itself runs on the oracle here (no GPU in this container); what is under test
is tantivy_b200/sharding.py — the same code bench.py runs over NCCL."""
import os
import socket

import numpy as np
import pytest

torch = pytest.importorskip("torch")
import torch.distributed as dist  # noqa: E402
import torch.multiprocessing as mp  # noqa: E402

import tantivy_b200 as T  # noqa: E402
from tantivy_b200._abi import TQ_OP_AND, TQ_OP_OR, TQ_OP_TERM  # noqa: E402
from tantivy_b200.sharding import ShardedIndex, assign_segments, merge_rows_host  # noqa: E402

DENS = [0.2, 0.05, 0.01, 0.001]
N_SEG, DOCS = 4, 60_000
QUERIES = [(TQ_OP_TERM, [1], 10), (TQ_OP_AND, [0, 1], 10), (TQ_OP_OR, [0, 2, 3], 20), (TQ_OP_OR, [3, 1], 5), (TQ_OP_AND, [1, 2], 7)]
KMAX = 20


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    from oracle import tq_oracle as O
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ords = assign_segments(N_SEG, world, rank)
        ix = T.SynthIndex(len(ords), DOCS, DENS, seed=99, segment_base=rank, segment_stride=world, n_threads=2)
        shard = ShardedIndex(ix, ords, len(DENS), dist)
        oi = O.OracleIndex()
        shard.register(oi)
        qb = shard.marshal(QUERIES)
        sc, sg, dc, ct = oi.search_batch(qb, mode=0)
        nq = len(QUERIES)
        gathered = []
        for arr, dt in ((sc, torch.float32), (sg.astype(np.int64), torch.int64), (dc.astype(np.int64), torch.int64), (ct.astype(np.int64), torch.int64)):
            t = torch.from_numpy(np.ascontiguousarray(arr)).to(dt)
            lst = [torch.empty_like(t) for _ in range(world)]
            dist.all_gather(lst, t)
            gathered.append(torch.stack(lst).numpy())
        merged = merge_rows_host(gathered[0], gathered[1].astype(np.uint32), gathered[2].astype(np.uint32), gathered[3].astype(np.uint32), KMAX)
        if rank == 0:
            out.put(dict(df=shard.df.tolist(), docs=shard.total_docs, tokens=shard.total_tokens, avg=float(shard.avg),
                         merged=[m.tolist() for m in merged]))
    finally:
        dist.destroy_process_group()


def test_two_rank_sharded_search_matches_single_process():
    from oracle import tq_oracle as O
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    res = out.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # single process over all four segments
    ix = T.SynthIndex(N_SEG, DOCS, DENS, seed=99, n_threads=2)
    shard = ShardedIndex(ix, list(range(N_SEG)), len(DENS))
    assert res["df"] == shard.df.tolist() and res["docs"] == shard.total_docs and res["tokens"] == shard.total_tokens
    assert res["avg"] == float(shard.avg)
    oi = O.OracleIndex()
    shard.register(oi)
    sc, sg, dc, ct = oi.search_batch(shard.marshal(QUERIES), mode=0)
    m_sc, m_sg, m_dc, m_ct = [np.array(x) for x in res["merged"]]
    for q, (_, _, k) in enumerate(QUERIES):
        n = int(ct[q])
        assert n <= k and int(m_ct[q]) >= n  # the merge ran with k = KMAX for every query
        assert (m_sg[q, :n] == sg[q, :n]).all() and (m_dc[q, :n] == dc[q, :n]).all()
        assert (m_sc[q, :n].astype(np.float32) == sc[q, :n]).all()


def test_assign_segments_partitions():
    for world in (1, 2, 3, 4, 8):
        seen = sorted(s for r in range(world) for s in assign_segments(8, world, r))
        assert seen == list(range(8))
