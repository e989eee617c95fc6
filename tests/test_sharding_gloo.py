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
from tantivy_b200.sharding import (ShardedIndex, assign_parts, assign_segments, doc_range, range_alive_bitset, exchange_thresholds, key_scores, kth_of_gathered_keys,  # noqa: E402
                                   local_topkeys, merge_rows_host, score_keys)

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


def _split_worker(rank, world, port, out, n_seg, docs):
    """SURVEY.md §8(e), second half: fewer segments than ranks -> every segment is split by doc-id range (assign_parts); a part
    behaves like a segment of its own (here on the oracle: the range as an alive bitset), statistics count the segment once."""
    from oracle import tq_oracle as O
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        units = assign_parts(n_seg, world, rank)
        ords = [u[0] for u in units]
        ix = T.SynthIndex(len(ords), docs, DENS, seed=99, segment_base=ords[0], segment_stride=n_seg, n_threads=2)
        shard = ShardedIndex(ix, ords, len(DENS), dist, parts=[(u[1], u[2]) for u in units])
        oi = O.OracleIndex()
        shard.register(oi)
        sc, sg, dc, ct = oi.search_batch(shard.marshal(QUERIES), mode=0)
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


def test_one_segment_split_by_doc_range_over_two_ranks():
    from oracle import tq_oracle as O
    docs = 150_001  # not a multiple of the tile: the last part is the short one
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_split_worker, args=(r, 2, port, out, 1, docs)) for r in range(2)]
    for p in procs:
        p.start()
    res = out.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    ix = T.SynthIndex(1, docs, DENS, seed=99, n_threads=2)
    shard = ShardedIndex(ix, [0], len(DENS))
    assert res["df"] == shard.df.tolist() and res["docs"] == shard.total_docs and res["tokens"] == shard.total_tokens
    assert res["avg"] == float(shard.avg)
    oi = O.OracleIndex()
    shard.register(oi)
    sc, sg, dc, ct = oi.search_batch(shard.marshal(QUERIES), mode=0)
    m_sc, m_sg, m_dc, m_ct = [np.array(x) for x in res["merged"]]
    for q, (_, _, k) in enumerate(QUERIES):
        n = int(ct[q])
        assert n == k and int(m_ct[q]) >= n
        assert (m_sg[q, :n] == sg[q, :n]).all() and (m_dc[q, :n] == dc[q, :n]).all()
        assert (m_sc[q, :n].astype(np.float32) == sc[q, :n]).all()


def test_assign_parts_and_doc_ranges_partition():
    for n_seg, world in ((8, 1), (8, 2), (8, 8), (3, 2), (1, 2), (1, 8), (2, 8), (3, 8)):
        units = [u for r in range(world) for u in assign_parts(n_seg, world, r)]
        for s in range(n_seg):
            mine = sorted((u[1], u[2]) for u in units if u[0] == s)
            assert mine and mine == [(p, mine[0][1]) for p in range(mine[0][1])]  # every part of every segment exactly once
    for max_doc in (0, 1, 1023, 1024, 1025, 150_001, 10_000_000):
        for n_parts in (1, 2, 3, 8):
            cuts = [doc_range(max_doc, p, n_parts) for p in range(n_parts)]
            assert cuts[0][0] == 0 and cuts[-1][1] == max_doc
            assert all(cuts[i][1] == cuts[i + 1][0] for i in range(n_parts - 1))
            assert all(lo % 1024 == 0 for lo, _ in cuts if lo < max_doc)
    bits = range_alive_bitset(21, 3, 18)
    assert bits.tolist() == [0b11111000, 0b11111111, 0b00000011]


def test_assign_segments_partitions():
    for world in (1, 2, 3, 4, 8):
        seen = sorted(s for r in range(world) for s in assign_segments(8, world, r))
        assert seen == list(range(8))


def _threshold_worker(rank, world, port, out):
    """The cross-shard threshold protocol of CrossGpuMerger.run on CPU: every rank finds its local k-th best (here with
    the oracle over a SAMPLE of its shard: one segment), the keys are max-reduced over gloo, every rank then searches
    its whole shard with the exchanged bound as tq_query.threshold."""
    from oracle import tq_oracle as O
    from tantivy_b200._abi import QueryBatch
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ords = assign_segments(N_SEG, world, rank)
        ix = T.SynthIndex(len(ords), DOCS, DENS, seed=99, segment_base=rank, segment_stride=world, n_threads=2)
        shard = ShardedIndex(ix, ords, len(DENS), dist)
        oi = O.OracleIndex()
        shard.register(oi)
        qb = shard.marshal(QUERIES)
        nq = len(QUERIES)
        # "threshold round": the first local segment only
        sample = shard.marshal(QUERIES)
        keep = sample.term_segs["segment_ord"] == ords[0]
        for i in range(nq):  # restrict every query to the sampled segment
            row = sample.q[i]
            first = (int(row["term_segs"]) - sample.term_segs.ctypes.data) // sample.term_segs.itemsize
            sel = [j for j in range(first, first + int(row["n_term_segs"])) if keep[j]]
            sample.term_segs[first:first + len(sel)] = sample.term_segs[sel]
            row["n_term_segs"] = len(sel)
        sc, _, _, ct = oi.search_batch(sample, mode=0)
        ks = [q[2] for q in QUERIES]
        local = np.array([score_keys(sc[i, ks[i] - 1]) if ct[i] >= ks[i] else 0 for i in range(nq)], dtype=np.int64)
        keys = torch.from_numpy(local.copy())
        exchange_thresholds(dist, keys)
        glob = keys.numpy()
        assert (glob >= local).all()
        # main pass: everything strictly above the next lower score stays, i.e. every score >= the exchanged bound
        thr = [float(np.nextafter(key_scores(glob[i]), np.float32(-np.inf))) if glob[i] else None for i in range(nq)]
        queries = []
        for i, (op, terms, k) in enumerate(QUERIES):
            queries.append((op, terms, k))
        main = shard.marshal(queries)
        for i in range(nq):
            if thr[i] is not None:
                main.q[i]["flags"] = 1
                main.q[i]["threshold"] = thr[i]
        rows_without = int(oi.search_batch(qb, mode=0)[3].sum())  # the same shard without the exchanged bounds
        sc, sg, dc, ct = oi.search_batch(main, mode=0)
        rows_with = torch.tensor([int(ct.sum()), rows_without])
        dist.all_reduce(rows_with)
        gathered = []
        for arr, dt in ((sc, torch.float32), (sg.astype(np.int64), torch.int64), (dc.astype(np.int64), torch.int64), (ct.astype(np.int64), torch.int64)):
            t = torch.from_numpy(np.ascontiguousarray(arr)).to(dt)
            lst = [torch.empty_like(t) for _ in range(world)]
            dist.all_gather(lst, t)
            gathered.append(torch.stack(lst).numpy())
        merged = merge_rows_host(gathered[0], gathered[1].astype(np.uint32), gathered[2].astype(np.uint32), gathered[3].astype(np.uint32), KMAX)
        if rank == 0:
            out.put(dict(merged=[m.tolist() for m in merged], rows=rows_with.tolist(), bounds=glob.tolist()))
    finally:
        dist.destroy_process_group()


def test_two_rank_threshold_exchange_loses_nothing():
    from oracle import tq_oracle as O
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_threshold_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    res = out.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    ix = T.SynthIndex(N_SEG, DOCS, DENS, seed=99, n_threads=2)
    shard = ShardedIndex(ix, list(range(N_SEG)), len(DENS))
    oi = O.OracleIndex()
    shard.register(oi)
    sc, sg, dc, ct = oi.search_batch(shard.marshal(QUERIES), mode=0)
    m_sc, m_sg, m_dc, m_ct = [np.array(x) for x in res["merged"]]
    for i, (_, _, k) in enumerate(QUERIES):
        n = int(ct[i])
        assert int(m_ct[i]) >= n  # (the host merge keeps up to KMAX rows; the first k are the answer)
        assert (m_sg[i, :n] == sg[i, :n]).all() and (m_dc[i, :n] == dc[i, :n]).all()
        assert (m_sc[i, :n].astype(np.float32) == sc[i, :n]).all()
    assert any(b > 0 for b in res["bounds"])  # the exchange did hand out bounds
    assert res["rows"][0] <= res["rows"][1]     # ... which can only remove rows
    for i, (_, _, k) in enumerate(QUERIES):    # every exchanged bound is a lower bound of the global k-th best score
        if res["bounds"][i] and int(ct[i]) == k:
            assert key_scores(res["bounds"][i]) <= sc[i, k - 1]
    # key mapping round trip
    x = np.array([0.0, 3.25e-3, 1.5, 1e9], dtype=np.float32)
    assert (key_scores(score_keys(x)) == x).all() and (np.diff(score_keys(x)) > 0).all()


def _topkeys_worker(rank, world, port, out):
    """The exact exchange of CrossGpuMerger.run on CPU: every rank exports the k best score keys of a SAMPLE of its shard (one
    segment), one all-gather moves them, the k-th best key of the union is every rank's bound (k_topkeys_export /
    k_theta_from_keys; host twins local_topkeys / kth_of_gathered_keys)."""
    from oracle import tq_oracle as O
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ords = assign_segments(N_SEG, world, rank)
        ix = T.SynthIndex(len(ords), DOCS, DENS, seed=99, segment_base=rank, segment_stride=world, n_threads=2)
        shard = ShardedIndex(ix, ords, len(DENS), dist)
        oi = O.OracleIndex()
        shard.register(oi)
        nq = len(QUERIES)
        ks = [q[2] for q in QUERIES]
        sample = shard.marshal(QUERIES)
        keep = sample.term_segs["segment_ord"] == ords[0]
        for i in range(nq):
            row = sample.q[i]
            first = (int(row["term_segs"]) - sample.term_segs.ctypes.data) // sample.term_segs.itemsize
            sel = [j for j in range(first, first + int(row["n_term_segs"])) if keep[j]]
            sample.term_segs[first:first + len(sel)] = sample.term_segs[sel]
            row["n_term_segs"] = len(sel)
        sc, _, _, ct = oi.search_batch(sample, mode=0)
        mine = torch.from_numpy(local_topkeys(sc, ct, ks, KMAX))
        lst = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(lst, mine)
        gathered = torch.stack(lst).numpy()
        bound = kth_of_gathered_keys(gathered, ks)
        # the old protocol's bound: MAX over ranks of each rank's own k-th best
        own = np.array([score_keys(sc[i, ks[i] - 1]) if ct[i] >= ks[i] else 0 for i in range(nq)], dtype=np.int64)
        t = torch.from_numpy(own.copy())
        exchange_thresholds(dist, t)
        if rank == 0:
            out.put(dict(bound=bound.tolist(), max_of_own=t.numpy().tolist(), gathered=gathered.tolist()))
    finally:
        dist.destroy_process_group()


def test_two_rank_topkeys_exchange_gives_the_union_kth():
    from oracle import tq_oracle as O
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_topkeys_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    res = out.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # single process over the two sampled segments (rank r sampled its first segment: global ordinals 0 and 1)
    ix = T.SynthIndex(N_SEG, DOCS, DENS, seed=99, n_threads=2)
    shard = ShardedIndex(ix, list(range(N_SEG)), len(DENS))
    oi = O.OracleIndex()
    shard.register(oi)
    qb = shard.marshal(QUERIES)
    keep = (qb.term_segs["segment_ord"] == 0) | (qb.term_segs["segment_ord"] == 1)
    for i in range(len(QUERIES)):
        row = qb.q[i]
        first = (int(row["term_segs"]) - qb.term_segs.ctypes.data) // qb.term_segs.itemsize
        sel = [j for j in range(first, first + int(row["n_term_segs"])) if keep[j]]
        qb.term_segs[first:first + len(sel)] = qb.term_segs[sel]
        row["n_term_segs"] = len(sel)
    sc, _, _, ct = oi.search_batch(qb, mode=0)
    full_sc, _, _, full_ct = oi.search_batch(shard.marshal(QUERIES), mode=0)
    for i, (_, _, k) in enumerate(QUERIES):
        want = int(score_keys(sc[i, k - 1])) if ct[i] >= k else 0
        assert res["bound"][i] == want              # exactly the k-th best one process finds over the same docs
        assert res["bound"][i] >= res["max_of_own"][i]  # never looser than the all-reduce(MAX) of the ranks' own k-th bests
        if full_ct[i] >= k:                         # and a valid lower bound of the final k-th score
            assert res["bound"][i] <= int(score_keys(full_sc[i, k - 1]))
