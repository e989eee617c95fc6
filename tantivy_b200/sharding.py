"""Segment sharding across GPUs: one process per GPU, segments round-robin over ranks, BM25
statistics summed over all shards, per-rank top-k rows all-gathered and merged.

This is the multi-GPU form of the reference's only parallel axis, one task per segment
(src/core/executor.rs:60-100, src/core/searcher.rs:229-235): the per-segment fruits of a rank are
already merged on its GPU (k_final); the cross-rank step is merge_fruits again
(src/collector/sort_key_top_collector.rs:54-60,76-95) over `world` lists of k rows.
The statistics follow Bm25StatisticsProvider for Searcher (src/query/bm25.rs:27-50):
N = sum of max_doc, avg fieldnorm = sum(total_num_tokens) / N, n = sum of doc_freq."""
import numpy as np

from ._abi import QueryBatch
from .lib import bm25_weight


def assign_segments(n_segments, world, rank):
    """Global segment ordinals owned by `rank` (round-robin: segments are equally sized)."""
    return [s for s in range(n_segments) if s % world == rank]


def global_statistics(local_doc_freq, local_tokens, local_docs, dist=None, device=None):
    """Sums per-term doc_freq, token count and doc count over all ranks (all_reduce)."""
    stats = np.concatenate([np.asarray(local_doc_freq, dtype=np.int64), [int(local_tokens), int(local_docs)]]).astype(np.int64)
    if dist is not None and dist.is_initialized() and dist.get_world_size() > 1:
        import torch
        t = torch.from_numpy(stats)
        if device is not None:
            t = t.to(device)
        dist.all_reduce(t)
        stats = t.cpu().numpy()
    nt = len(local_doc_freq)
    df = stats[:nt]
    tokens, docs = int(stats[nt]), int(stats[nt + 1])
    avg = np.float32(np.float32(tokens) / np.float32(docs)) if docs else np.float32(0)
    return df, tokens, docs, avg


def merge_rows_host(scores, segs, docs, counts, k):
    """Host reference of the cross-rank merge: inputs [world, nq, stride] / [world, nq]; returns rows
    ordered (score desc, segment_ord asc, doc asc), like TopDocs (top_score_collector.rs:591-600)."""
    world, nq, stride = scores.shape
    o_sc = np.zeros((nq, k), np.float32)
    o_sg = np.zeros((nq, k), np.uint32)
    o_dc = np.zeros((nq, k), np.uint32)
    o_ct = np.zeros(nq, np.uint32)
    for q in range(nq):
        rows = []
        for r in range(world):
            n = min(int(counts[r, q]), stride)
            rows += [(-float(scores[r, q, i]), int(segs[r, q, i]), int(docs[r, q, i])) for i in range(n)]
        rows.sort()
        rows = rows[:k]
        o_ct[q] = len(rows)
        for i, (ns, g, d) in enumerate(rows):
            o_sc[q, i], o_sg[q, i], o_dc[q, i] = -ns, g, d
    return o_sc, o_sg, o_dc, o_ct


def score_keys(scores):
    """Order-preserving u32 image of f32 scores, as int64 (host twin of score_to_key in csrc/tq_device.cuh):
    the unit in which shards exchange thresholds; 0 = no bound."""
    u = np.asarray(scores, dtype=np.float32).view(np.uint32).astype(np.int64)
    return np.where(u >> 31, u ^ 0xFFFFFFFF, u ^ 0x80000000)


def key_scores(keys):
    """Inverse of score_keys."""
    k = np.asarray(keys, dtype=np.int64)
    u = np.where(k >> 31, k ^ 0x80000000, k ^ 0xFFFFFFFF).astype(np.uint32)
    return u.view(np.float32)


def exchange_thresholds(dist, keys):
    """Element-wise MAX of every rank's per-query threshold keys (a torch int64 tensor, in place): any rank's k-th best
    score is a lower bound of the global k-th best, so the largest one is the bound every rank may prune against."""
    dist.all_reduce(keys, op=dist.ReduceOp.MAX)
    return keys


class ShardedIndex:
    """This rank's segments of a SynthIndex-like index plus the global statistics.

    `index` must expose n_segments, max_doc[], total_num_tokens[], term_info[s][t], record_option,
    body(s), fieldnorm(s); `global_ords[i]` is the global ordinal of local segment i."""

    def __init__(self, index, global_ords, n_terms, dist=None, device=None):
        self.ix = index
        self.global_ords = list(global_ords)
        df = np.zeros(n_terms, dtype=np.int64)
        tokens = docs = 0
        if index is not None:
            for s in range(index.n_segments):
                df += np.array([index.term_info[s][t][0] for t in range(n_terms)], dtype=np.int64)
            tokens, docs = sum(index.total_num_tokens), sum(index.max_doc)
        self.df, self.total_tokens, self.total_docs, self.avg = global_statistics(df, tokens, docs, dist, device)
        self.index_bytes = sum(index.body(s).size + index.fieldnorm(s).size for s in range(index.n_segments)) if index is not None else 0

    def register(self, target):
        for i, g in enumerate(self.global_ords):
            target.segment_register(g, 0, self.ix.max_doc[i], self.ix.record_option, self.ix.body(i), self.ix.fieldnorm(i), None)

    def marshal(self, queries):
        """queries: iterable of (op code, [term ordinals], k) -> QueryBatch restricted to this rank's segments."""
        out = []
        for op, terms, k in queries:
            weights = [bm25_weight(int(self.df[t]), self.total_docs, 1.0) for t in terms]
            term_segs = []
            for clause, t in enumerate(terms):
                for i, g in enumerate(self.global_ords):
                    d, st, en = self.ix.term_info[i][t]
                    if d:
                        term_segs.append((clause, g, 0, d, st, en))
            out.append(dict(op=op, k=k, weights=weights, avg_fieldnorm=[self.avg] * len(terms), term_segs=term_segs))
        return QueryBatch(out)


class CrossGpuMerger:
    """NCCL all-gather of every rank's result rows + device merge (K7)."""

    def __init__(self, ctx, dist, device, nq, k):
        import torch
        self.ctx, self.dist, self.nq, self.k = ctx, dist, nq, k
        self.world = dist.get_world_size()
        f32, i32 = torch.float32, torch.int32
        mk = lambda shape, dt: torch.empty(shape, dtype=dt, device=device)  # noqa: E731
        self.l = (mk((nq, k), f32), mk((nq, k), i32), mk((nq, k), i32), mk((nq,), i32))
        self.g = (mk((self.world, nq, k), f32), mk((self.world, nq, k), i32), mk((self.world, nq, k), i32), mk((self.world, nq), i32))
        self.o = (mk((nq, k), f32), mk((nq, k), i32), mk((nq, k), i32), mk((nq,), i32))
        self.theta = mk((nq,), torch.int64)
        self.torch = torch

    def run(self, batch):
        """Runs a prepared batch with the cross-rank threshold exchange: after every threshold round of the unions the
        per-query score keys are all-reduced (MAX, 8 bytes per query) over NCCL, so that every rank prunes against the
        best lower bound any rank has found so far."""
        n = batch.phases()
        for phase in range(n):
            batch.run_phase(phase)
            if phase + 1 < n:
                batch.thresholds_export_dev(self.theta.data_ptr())  # waits for the batch's stream
                exchange_thresholds(self.dist, self.theta)
                self.torch.cuda.synchronize()
                batch.thresholds_import_dev(self.theta.data_ptr())

    def __call__(self, batch):
        """batch: a finished tantivy_b200.Batch of this rank. Returns merged device tensors (every rank)."""
        batch.results_copy_dev(*[t.data_ptr() for t in self.l])  # waits for the batch's stream
        for g, l in zip(self.g, self.l):
            self.dist.all_gather_into_tensor(g, l)
        self.torch.cuda.synchronize()
        self.ctx.merge_topk_dev(self.world, self.nq, self.k, self.k, *[t.data_ptr() for t in self.g], *[t.data_ptr() for t in self.o])
        return self.o
