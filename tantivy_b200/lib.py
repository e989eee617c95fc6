"""ctypes binding of the product library tantivy_b200/_lib/libtantivy_b200.so (the C ABI of
include/tantivy_b200.h).  The library is CUDA-only: if it is missing, importing fails loudly;
if no GPU is present, tq_ctx_create fails loudly.  Nothing here falls back to a CPU path."""
import ctypes as C
import os

import numpy as np

from ._abi import Query, QueryBatch, Stats, TermSeg, f32p, ptr, u8p, u32p, u64p

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.environ.get("TANTIVY_B200_LIB") or os.path.join(_HERE, "_lib", "libtantivy_b200.so")  # (override: A/B builds)


class TqError(RuntimeError):
    pass


def _load():
    if not os.path.exists(SO_PATH):
        raise ImportError(
            f"{SO_PATH} is missing: build it with `make -C tantivy_b200/csrc` (or __graft_entry__.build()). "
            "There is no CPU fallback for the query path.")
    lib = C.CDLL(SO_PATH)
    vp, sz = C.c_void_p, C.c_size_t
    lib.tq_last_error.restype = C.c_char_p
    lib.tq_last_error.argtypes = [vp]
    lib.tq_ctx_create.argtypes = [C.c_int, C.POINTER(vp)]
    lib.tq_ctx_destroy.argtypes = [vp]
    lib.tq_get_stats.argtypes = [vp, C.POINTER(Stats)]
    lib.tq_segment_register.argtypes = [vp, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int, u8p, sz, u8p, sz, u8p, sz]
    lib.tq_segment_register_positions.argtypes = [vp, C.c_uint32, C.c_uint32, u8p, sz]
    lib.tq_segment_unregister.argtypes = [vp, C.c_uint32, C.c_uint32]
    lib.tq_segment_set_doc_range.argtypes = [vp, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32]
    lib.tq_search_batch.argtypes = [vp, C.POINTER(Query), sz, C.c_uint32, f32p, u32p, u32p, u32p]
    lib.tq_batch_prepare.argtypes = [vp, C.POINTER(Query), sz, C.POINTER(vp)]
    lib.tq_count_batch.argtypes = [vp, C.POINTER(Query), sz, u64p]
    lib.tq_batch_run.argtypes = [vp]
    lib.tq_batch_run_phase.argtypes = [vp, C.c_int]
    lib.tq_batch_phases.argtypes = [vp]
    lib.tq_batch_thresholds_export_dev.argtypes = [vp, vp]
    lib.tq_batch_thresholds_import_dev.argtypes = [vp, vp]
    lib.tq_batch_stream.argtypes = [vp, C.POINTER(vp)]
    lib.tq_batch_topkeys_export_dev.argtypes = [vp, vp, C.c_uint32]
    lib.tq_batch_thresholds_from_keys_dev.argtypes = [vp, vp, C.c_uint32, C.c_uint32]
    lib.tq_batch_fetch.argtypes = [vp, C.c_uint32, f32p, u32p, u32p, u32p]
    lib.tq_batch_results_dev.argtypes = [vp, C.POINTER(vp), C.POINTER(vp), C.POINTER(vp), C.POINTER(vp), C.POINTER(C.c_uint32)]
    lib.tq_batch_results_copy_dev.argtypes = [vp, vp, vp, vp, vp]
    lib.tq_batch_destroy.argtypes = [vp]
    lib.tq_merge_topk_dev.argtypes = [vp, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, vp, vp, vp, vp, vp, vp, vp, vp]
    lib.tq_batch_results_pack_dev.argtypes = [vp, vp]
    lib.tq_batch_results_pack_dev_async.argtypes = [vp, vp]
    lib.tq_merge_topk_packed_dev.argtypes = [vp, vp, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, vp, sz, vp]
    lib.tq_multi_create.argtypes = [C.POINTER(C.c_int), C.c_int, C.POINTER(vp)]
    lib.tq_multi_destroy.argtypes = [vp]
    lib.tq_multi_last_error.restype = C.c_char_p
    lib.tq_multi_last_error.argtypes = [vp]
    lib.tq_multi_num_devices.argtypes = [vp]
    lib.tq_multi_segment_register.argtypes = [vp, C.c_int, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int, u8p, sz, u8p, sz, u8p, sz]
    lib.tq_multi_segment_register_split.argtypes = [vp, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int, u8p, sz, u8p, sz, u8p, sz]
    lib.tq_multi_search_batch.argtypes = [vp, C.POINTER(Query), sz, C.c_uint32, f32p, u32p, u32p, u32p]
    lib.tq_decode_postings.argtypes = [vp, C.POINTER(TermSeg), u32p, u32p]
    lib.tq_block_table.argtypes = [vp, C.POINTER(TermSeg), C.c_float, C.c_float, u32p, f32p]
    lib.tq_bm25_idf.restype = C.c_float
    lib.tq_bm25_idf.argtypes = [C.c_uint64, C.c_uint64]
    lib.tq_bm25_weight.restype = C.c_float
    lib.tq_bm25_weight.argtypes = [C.c_uint64, C.c_uint64, C.c_float]
    lib.tq_bm25_tf_cache.argtypes = [C.c_float, f32p]
    lib.tq_id_to_fieldnorm.restype = C.c_uint32
    lib.tq_id_to_fieldnorm.argtypes = [C.c_uint8]
    lib.tq_fieldnorm_to_id.restype = C.c_uint8
    lib.tq_fieldnorm_to_id.argtypes = [C.c_uint32]
    lib.tq_field_writer_create.argtypes = [C.c_int, C.c_uint64, u8p, C.c_uint32, C.POINTER(vp)]
    lib.tq_field_writer_add_term.argtypes = [vp, u32p, u32p, C.c_uint32, u64p, u64p]
    lib.tq_field_writer_body.argtypes = [vp, C.POINTER(u8p), C.POINTER(sz)]
    lib.tq_field_writer_destroy.argtypes = [vp]
    # synthetic segment generator (csrc/synth.cpp)
    lib.tqs_generate.restype = vp
    lib.tqs_generate.argtypes = [C.c_uint32, C.c_uint32, C.POINTER(C.c_double), C.c_uint32, C.c_uint64, C.c_int, C.c_int, C.c_uint32, C.c_uint32]
    lib.tqs_destroy.argtypes = [vp]
    lib.tqs_num_segments.restype = C.c_uint32
    lib.tqs_num_segments.argtypes = [vp]
    lib.tqs_max_doc.restype = C.c_uint32
    lib.tqs_max_doc.argtypes = [vp, C.c_uint32]
    lib.tqs_total_num_tokens.restype = C.c_uint64
    lib.tqs_total_num_tokens.argtypes = [vp, C.c_uint32]
    lib.tqs_body.argtypes = [vp, C.c_uint32, C.POINTER(u8p), C.POINTER(sz)]
    lib.tqs_fieldnorm.argtypes = [vp, C.c_uint32, C.POINTER(u8p), C.POINTER(sz)]
    lib.tqs_term_info.argtypes = [vp, C.c_uint32, C.c_uint32, u32p, u64p, u64p]
    lib.tqs_positions.argtypes = [vp, C.c_uint32, C.POINTER(u8p), C.POINTER(sz)]
    lib.tqs_term_pos.argtypes = [vp, C.c_uint32, C.c_uint32, u64p, u64p]
    return lib


LIB = _load()


def _check(rc, ctx=None):
    if rc != 0:
        raise TqError(f"tantivy_b200 error {rc}: {LIB.tq_last_error(ctx).decode(errors='replace')}")


def bm25_idf(doc_freq, doc_count):
    return LIB.tq_bm25_idf(doc_freq, doc_count)


def bm25_weight(doc_freq, doc_count, boost=1.0):
    return LIB.tq_bm25_weight(doc_freq, doc_count, boost)


def bm25_tf_cache(avg):
    out = np.zeros(256, dtype=np.float32)
    LIB.tq_bm25_tf_cache(avg, ptr(out, f32p))
    return out


def id_to_fieldnorm(i):
    return LIB.tq_id_to_fieldnorm(i)


def fieldnorm_to_id(f):
    return LIB.tq_fieldnorm_to_id(f)


class FieldWriter:
    """Host-side PostingsSerializer equivalent (src/postings/serializer.rs:353-481)."""

    def __init__(self, record_option, total_num_tokens, fieldnorm_ids, max_doc):
        self._fn = None if fieldnorm_ids is None else np.ascontiguousarray(fieldnorm_ids, dtype=np.uint8)
        h = C.c_void_p()
        _check(LIB.tq_field_writer_create(record_option, int(total_num_tokens), ptr(self._fn, u8p), max_doc, C.byref(h)))
        self.h = h

    def add_term(self, docs, tfs=None):
        docs = np.ascontiguousarray(docs, dtype=np.uint32)
        tfs = None if tfs is None else np.ascontiguousarray(tfs, dtype=np.uint32)
        s, e = C.c_uint64(), C.c_uint64()
        _check(LIB.tq_field_writer_add_term(self.h, ptr(docs, u32p), ptr(tfs, u32p), len(docs), C.byref(s), C.byref(e)))
        return int(s.value), int(e.value)

    def body(self):
        p, n = u8p(), C.c_size_t()
        _check(LIB.tq_field_writer_body(self.h, C.byref(p), C.byref(n)))
        return np.ctypeslib.as_array(p, shape=(n.value,)).copy()

    def __del__(self):
        if getattr(self, "h", None):
            LIB.tq_field_writer_destroy(self.h)
            self.h = None


class SynthIndex:
    """Deterministic synthetic segments in tantivy format (csrc/synth.cpp)."""

    def __init__(self, n_segments, docs_per_segment, densities, seed=0x7A6E7469, record_option=1, n_threads=0, segment_base=0,
                 segment_stride=1):
        dens = np.ascontiguousarray(densities, dtype=np.float64)
        self.densities = dens
        self.n_threads = n_threads or (os.cpu_count() or 1)
        self.h = LIB.tqs_generate(n_segments, docs_per_segment, dens.ctypes.data_as(C.POINTER(C.c_double)), len(dens), seed,
                                  record_option, self.n_threads, segment_base, segment_stride)
        self.n_segments = n_segments
        self.record_option = record_option
        self.max_doc = [LIB.tqs_max_doc(self.h, s) for s in range(n_segments)]
        self.total_num_tokens = [LIB.tqs_total_num_tokens(self.h, s) for s in range(n_segments)]
        # term_info[s][t] = (doc_freq, start, end); term_pos[s][t] = (positions_start, positions_end) with record_option 2
        self.term_info, self.term_pos = [], []
        df, st, en = C.c_uint32(), C.c_uint64(), C.c_uint64()
        for s in range(n_segments):
            row, prow = [], []
            for t in range(len(dens)):
                LIB.tqs_term_info(self.h, s, t, C.byref(df), C.byref(st), C.byref(en))
                row.append((df.value, st.value, en.value))
                if record_option == 2:
                    LIB.tqs_term_pos(self.h, s, t, C.byref(st), C.byref(en))
                    prow.append((st.value, en.value))
            self.term_info.append(row)
            self.term_pos.append(prow)

    def body(self, s):
        p, n = u8p(), C.c_size_t()
        LIB.tqs_body(self.h, s, C.byref(p), C.byref(n))
        return np.ctypeslib.as_array(p, shape=(n.value,))

    def fieldnorm(self, s):
        p, n = u8p(), C.c_size_t()
        LIB.tqs_fieldnorm(self.h, s, C.byref(p), C.byref(n))
        return np.ctypeslib.as_array(p, shape=(n.value,))

    def positions(self, s):
        """The segment's `.pos` body (record_option 2)."""
        p, n = u8p(), C.c_size_t()
        LIB.tqs_positions(self.h, s, C.byref(p), C.byref(n))
        return np.ctypeslib.as_array(p, shape=(max(n.value, 1),))[:n.value]

    def num_docs(self):
        return sum(self.max_doc)

    def avg_fieldnorm(self):
        return np.float32(np.float32(sum(self.total_num_tokens)) / np.float32(self.num_docs()))

    def doc_freq(self, t):
        return sum(self.term_info[s][t][0] for s in range(self.n_segments))

    def register(self, target, field=0, segment_base=0):
        """target: Context (GPU) or an oracle index with the same segment_register signature."""
        for s in range(self.n_segments):
            target.segment_register(segment_base + s, field, self.max_doc[s], self.record_option, self.body(s), self.fieldnorm(s), None)
            if self.record_option == 2:
                target.register_positions(segment_base + s, field, self.positions(s))

    def phrase_query(self, terms, k, field=0, segment_base=0, segments=None):
        """PhraseQuery over `terms` (in phrase order): ONE Bm25Weight::for_terms weight (the idfs add up, bm25.rs:95-129)."""
        n_docs = self.num_docs()
        avg = self.avg_fieldnorm()
        idf_sum = np.float32(0)
        for t in terms:
            idf_sum = np.float32(idf_sum + np.float32(bm25_idf(self.doc_freq(t), n_docs)))
        weight = np.float32(idf_sum * np.float32(2.2))
        term_segs, term_pos = [], []
        for clause, t in enumerate(terms):
            for s in (range(self.n_segments) if segments is None else segments):
                df, st, en = self.term_info[s][t]
                if df:
                    term_segs.append((clause, segment_base + s, field, df, st, en))
                    term_pos.append(self.term_pos[s][t])
        return dict(op=3, k=k, weights=[weight] * len(terms), avg_fieldnorm=[avg] * len(terms), term_segs=term_segs, term_pos=term_pos,
                    term_offset=list(range(len(terms))))

    def query(self, op, terms, k, field=0, segment_base=0, boost=1.0, segments=None):
        n_docs = self.num_docs()
        avg = self.avg_fieldnorm()
        weights, term_segs = [], []
        for clause, t in enumerate(terms):
            weights.append(bm25_weight(self.doc_freq(t), n_docs, boost))
            for s in (range(self.n_segments) if segments is None else segments):
                df, st, en = self.term_info[s][t]
                if df:
                    term_segs.append((clause, segment_base + s, field, df, st, en))
        return dict(op=op, k=k, weights=weights, avg_fieldnorm=[avg] * len(terms), term_segs=term_segs)

    def __del__(self):
        if getattr(self, "h", None):
            LIB.tqs_destroy(self.h)
            self.h = None


class Context:
    """tq_ctx: one CUDA device holding registered segments and the per-term block-table cache."""

    def __init__(self, device=0):
        h = C.c_void_p()
        _check(LIB.tq_ctx_create(device, C.byref(h)))
        self.h = h
        self.device = device

    def segment_register(self, segment_ord, field, max_doc, record_option, idx_body, fieldnorm=None, alive=None):
        idx_body = np.ascontiguousarray(idx_body, dtype=np.uint8)
        fn = None if fieldnorm is None else np.ascontiguousarray(fieldnorm, dtype=np.uint8)
        al = None if alive is None else np.ascontiguousarray(alive, dtype=np.uint8)
        _check(LIB.tq_segment_register(self.h, segment_ord, field, max_doc, record_option, ptr(idx_body, u8p), idx_body.size,
                                       ptr(fn, u8p), 0 if fn is None else fn.size, ptr(al, u8p), 0 if al is None else al.size), self.h)

    def register_positions(self, segment_ord, field, pos_bytes):
        """The field's `.pos` sub-file of a registered segment (phrase queries)."""
        p = np.ascontiguousarray(pos_bytes, dtype=np.uint8)
        _check(LIB.tq_segment_register_positions(self.h, segment_ord, field, ptr(p, u8p), p.size), self.h)

    def segment_unregister(self, segment_ord, field):
        _check(LIB.tq_segment_unregister(self.h, segment_ord, field), self.h)

    def segment_set_doc_range(self, segment_ord, field, doc_lo, doc_hi):
        """This context evaluates only the docs [doc_lo, doc_hi) of the segment (intra-segment split over GPUs)."""
        _check(LIB.tq_segment_set_doc_range(self.h, segment_ord, field, doc_lo, doc_hi), self.h)

    def search_batch(self, batch: QueryBatch, out=None):
        stride, scores, segs, docs, counts = out or batch.alloc_out()
        _check(LIB.tq_search_batch(self.h, batch.ptr, batch.nq, stride, ptr(scores, f32p), ptr(segs, u32p), ptr(docs, u32p),
                                   ptr(counts, u32p)), self.h)
        return scores, segs, docs, counts

    def count_batch(self, qb: QueryBatch):
        """Count collector: alive docs matching each query (tq_count_batch)."""
        out = np.zeros(max(qb.nq, 1), dtype=np.uint64)
        _check(LIB.tq_count_batch(self.h, qb.ptr, qb.nq, ptr(out, u64p)), self.h)
        return out[:qb.nq]

    def prepare(self, batch: QueryBatch):
        return Batch(self, batch)

    def decode_postings(self, term_seg):
        ts = TermSeg(*[int(x) for x in term_seg])
        docs = np.zeros(max(ts.doc_freq, 1), dtype=np.uint32)
        tfs = np.zeros(max(ts.doc_freq, 1), dtype=np.uint32)
        _check(LIB.tq_decode_postings(self.h, C.byref(ts), ptr(docs, u32p), ptr(tfs, u32p)), self.h)
        return docs[:ts.doc_freq], tfs[:ts.doc_freq]

    def block_table(self, term_seg, weight, avg_fieldnorm):
        ts = TermSeg(*[int(x) for x in term_seg])
        n = ts.doc_freq // 128
        last = np.zeros(max(n, 1), dtype=np.uint32)
        bm = np.zeros(max(n, 1), dtype=np.float32)
        _check(LIB.tq_block_table(self.h, C.byref(ts), weight, avg_fieldnorm, ptr(last, u32p), ptr(bm, f32p)), self.h)
        return last[:n], bm[:n]

    def stats(self):
        s = Stats()
        _check(LIB.tq_get_stats(self.h, C.byref(s)), self.h)
        return {name: (list(getattr(s, name)) if name in ("or_windows", "tile_counters") else getattr(s, name)) for name, _ in Stats._fields_}

    def merge_topk_dev(self, n_lists, nq, stride, k, scores, segs, docs, counts, out_scores, out_segs, out_docs, out_counts):
        """All arguments are raw device addresses (int), e.g. torch tensors' data_ptr()."""
        _check(LIB.tq_merge_topk_dev(self.h, n_lists, nq, stride, k, scores, segs, docs, counts, out_scores, out_segs, out_docs,
                                     out_counts), self.h)

    def merge_topk_packed_dev(self, stream, n_lists, nq, stride, k, packed, pitch_words, out_packed):
        """packed / out_packed: raw device addresses; stream: cudaStream_t as int (0 = the default stream).  No host sync."""
        _check(LIB.tq_merge_topk_packed_dev(self.h, stream, n_lists, nq, stride, k, packed, pitch_words, out_packed), self.h)

    def close(self):
        if getattr(self, "h", None):
            LIB.tq_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        self.close()


class MultiContext:
    """tq_multi: several devices behind one handle (segments sharded over them, in-process fan-out and merge)."""

    def __init__(self, devices):
        arr = (C.c_int * len(devices))(*devices)
        h = C.c_void_p()
        _check(LIB.tq_multi_create(arr, len(devices), C.byref(h)))
        self.h = h

    def segment_register(self, segment_ord, field, max_doc, record_option, idx_body, fieldnorm=None, alive=None, device_index=-1):
        idx_body = np.ascontiguousarray(idx_body, dtype=np.uint8)
        fn = None if fieldnorm is None else np.ascontiguousarray(fieldnorm, dtype=np.uint8)
        al = None if alive is None else np.ascontiguousarray(alive, dtype=np.uint8)
        rc = LIB.tq_multi_segment_register(self.h, device_index, segment_ord, field, max_doc, record_option, ptr(idx_body, u8p), idx_body.size,
                                           ptr(fn, u8p), 0 if fn is None else fn.size, ptr(al, u8p), 0 if al is None else al.size)
        if rc != 0:
            raise TqError(f"tantivy_b200 error {rc}: {LIB.tq_multi_last_error(self.h).decode(errors='replace')}")

    def segment_register_split(self, segment_ord, field, max_doc, record_option, idx_body, fieldnorm=None, alive=None):
        """One segment on every device of the handle, each evaluating its own doc range."""
        idx_body = np.ascontiguousarray(idx_body, dtype=np.uint8)
        fn = None if fieldnorm is None else np.ascontiguousarray(fieldnorm, dtype=np.uint8)
        al = None if alive is None else np.ascontiguousarray(alive, dtype=np.uint8)
        rc = LIB.tq_multi_segment_register_split(self.h, segment_ord, field, max_doc, record_option, ptr(idx_body, u8p), idx_body.size,
                                                 ptr(fn, u8p), 0 if fn is None else fn.size, ptr(al, u8p), 0 if al is None else al.size)
        if rc != 0:
            raise TqError(f"tantivy_b200 error {rc}: {LIB.tq_multi_last_error(self.h).decode(errors='replace')}")

    def search_batch(self, batch: QueryBatch, out=None):
        stride, scores, segs, docs, counts = out or batch.alloc_out()
        rc = LIB.tq_multi_search_batch(self.h, batch.ptr, batch.nq, stride, ptr(scores, f32p), ptr(segs, u32p), ptr(docs, u32p), ptr(counts, u32p))
        if rc != 0:
            raise TqError(f"tantivy_b200 error {rc}: {LIB.tq_multi_last_error(self.h).decode(errors='replace')}")
        return scores, segs, docs, counts

    def close(self):
        if getattr(self, "h", None):
            LIB.tq_multi_destroy(self.h)
            self.h = None

    def __del__(self):
        self.close()


class Batch:
    """tq_batch: prepare once, run (device only), fetch (D2H)."""

    def __init__(self, ctx: Context, qb: QueryBatch):
        self.ctx, self.qb = ctx, qb
        h = C.c_void_p()
        _check(LIB.tq_batch_prepare(ctx.h, qb.ptr, qb.nq, C.byref(h)), ctx.h)
        self.h = h

    def run(self):
        _check(LIB.tq_batch_run(self.h), self.ctx.h)

    def phases(self):
        return LIB.tq_batch_phases(self.h)

    def run_phase(self, phase):
        """Phases 0 .. phases()-1 in order; thresholds may be exchanged between two of them (tq_batch_run_phase)."""
        _check(LIB.tq_batch_run_phase(self.h, phase), self.ctx.h)

    def thresholds_export_dev(self, keys_dev):
        """keys_dev: device address of nq int64 (e.g. a torch tensor's data_ptr()); waits for the batch's stream."""
        _check(LIB.tq_batch_thresholds_export_dev(self.h, keys_dev), self.ctx.h)

    def thresholds_import_dev(self, keys_dev):
        _check(LIB.tq_batch_thresholds_import_dev(self.h, keys_dev), self.ctx.h)

    def stream(self):
        """The CUDA stream (cudaStream_t as int) every launch of this batch is enqueued on."""
        s = C.c_void_p()
        _check(LIB.tq_batch_stream(self.h, C.byref(s)), self.ctx.h)
        return s.value or 0

    def topkeys_export_dev(self, keys_dev, k_stride):
        """keys_dev: device address of nq * k_stride uint32; enqueued on the batch's stream, no host sync."""
        _check(LIB.tq_batch_topkeys_export_dev(self.h, keys_dev, k_stride), self.ctx.h)

    def thresholds_from_keys_dev(self, gathered_dev, n_shards, k_stride):
        _check(LIB.tq_batch_thresholds_from_keys_dev(self.h, gathered_dev, n_shards, k_stride), self.ctx.h)

    def fetch(self, out=None):
        stride, scores, segs, docs, counts = out or self.qb.alloc_out()
        _check(LIB.tq_batch_fetch(self.h, stride, ptr(scores, f32p), ptr(segs, u32p), ptr(docs, u32p), ptr(counts, u32p)), self.ctx.h)
        return scores, segs, docs, counts

    def results_dev(self):
        s, g, d, c = C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_void_p()
        stride = C.c_uint32()
        _check(LIB.tq_batch_results_dev(self.h, C.byref(s), C.byref(g), C.byref(d), C.byref(c), C.byref(stride)), self.ctx.h)
        return s.value, g.value, d.value, c.value, stride.value

    def results_copy_dev(self, scores, segs, docs, counts):
        """raw device addresses (e.g. torch tensor .data_ptr()) of [nq, kmax] / [nq] buffers"""
        _check(LIB.tq_batch_results_copy_dev(self.h, scores, segs, docs, counts), self.ctx.h)

    def results_pack_dev(self, packed):
        """packed: raw device address of 3 * nq * kmax + nq 32-bit words (scores | segment ords | docs | counts)."""
        _check(LIB.tq_batch_results_pack_dev(self.h, packed), self.ctx.h)

    def results_pack_dev_async(self, packed):
        """As results_pack_dev, enqueued behind the run WITHOUT waiting for it; packed holds 4 more words (the run's overflow
        flags: non-zero on any shard = take results_pack_dev)."""
        _check(LIB.tq_batch_results_pack_dev_async(self.h, packed), self.ctx.h)

    def close(self):
        if getattr(self, "h", None):
            LIB.tq_batch_destroy(self.h)
            self.h = None

    def __del__(self):
        self.close()
