"""ORACLE — TEST INFRASTRUCTURE ONLY.

ctypes binding of oracle/_build/libtq_oracle.so, the CPU restatement of the reference's
posting decode -> AND/OR -> BM25 -> top-k path.  Only tests/, __graft_entry__.smoke() and the
cpu_baseline / `--impl reference` legs of bench.py may import this module; the product
(tantivy_b200/) never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from tantivy_b200._abi import (Query, QueryBatch, TermSeg, f32p, ptr, u8p, u32p, u64p)

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libtq_oracle.so")


def build(force=False):
    if force or not os.path.exists(_SO):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return _SO


def _load():
    build()
    lib = C.CDLL(_SO)
    vp = C.c_void_p
    sz = C.c_size_t
    lib.tqo_index_create.restype = vp
    lib.tqo_index_destroy.argtypes = [vp]
    lib.tqo_last_error.restype = C.c_char_p
    lib.tqo_last_error.argtypes = [vp]
    lib.tqo_segment_register.argtypes = [vp, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int, u8p, sz, u8p, sz, u8p, sz]
    lib.tqo_search_batch.argtypes = [vp, C.POINTER(Query), sz, C.c_int, C.c_int, C.c_uint32, f32p, u32p, u32p, u32p]
    lib.tqo_count_batch.argtypes = [vp, C.POINTER(Query), sz, u64p]
    lib.tqo_segment_register_positions.argtypes = [vp, C.c_uint32, C.c_uint32, u8p, sz]
    lib.tqo_phrase_search.argtypes = [vp, vp, sz, C.c_uint32, C.c_float, C.c_float, C.c_uint32, sz, u32p, u32p, f32p, u32p, C.POINTER(sz)]
    lib.tqo_positions_serialize.restype = sz
    lib.tqo_positions_serialize.argtypes = [u32p, sz, sz, u8p, sz]
    lib.tqo_position_reader_open.restype = vp
    lib.tqo_position_reader_open.argtypes = [u8p, sz]
    lib.tqo_position_reader_read.argtypes = [vp, C.c_uint64, u32p, sz]
    lib.tqo_position_reader_close.argtypes = [vp]
    lib.tqo_decode_postings.argtypes = [vp, C.POINTER(TermSeg), u32p, u32p]
    lib.tqo_block_table.argtypes = [vp, C.POINTER(TermSeg), C.c_float, C.c_float, u32p, f32p]
    lib.tqo_term_scorer_open.restype = vp
    lib.tqo_term_scorer_open.argtypes = [vp, C.POINTER(TermSeg), C.c_float, C.c_float]
    lib.tqo_term_scorer_close.argtypes = [vp]
    for name in ("doc", "advance", "term_freq", "last_doc_in_block"):
        f = getattr(lib, "tqo_term_scorer_" + name)
        f.restype = C.c_uint32
        f.argtypes = [vp]
    lib.tqo_term_scorer_seek.restype = C.c_uint32
    lib.tqo_term_scorer_seek.argtypes = [vp, C.c_uint32]
    lib.tqo_term_scorer_seek_block.argtypes = [vp, C.c_uint32]
    for name in ("score", "block_max_score", "max_score"):
        f = getattr(lib, "tqo_term_scorer_" + name)
        f.restype = C.c_float
        f.argtypes = [vp]
    lib.tqo_bp4x_num_bits.restype = C.c_uint8
    lib.tqo_bp4x_num_bits.argtypes = [u32p]
    lib.tqo_bp4x_pack.argtypes = [u32p, C.c_uint8, u8p]
    lib.tqo_bp4x_unpack.argtypes = [u8p, C.c_uint8, u32p]
    lib.tqo_bp4x_unpack_scalar.argtypes = [u8p, C.c_uint8, u32p]
    lib.tqo_compress_block_sorted.restype = C.c_uint8
    lib.tqo_compress_block_sorted.argtypes = [u32p, C.c_uint32, u8p, C.POINTER(sz)]
    lib.tqo_compress_block_unsorted.restype = C.c_uint8
    lib.tqo_compress_block_unsorted.argtypes = [u32p, C.c_int, u8p, C.POINTER(sz)]
    lib.tqo_uncompress_block_sorted.restype = sz
    lib.tqo_uncompress_block_sorted.argtypes = [u8p, C.c_uint32, C.c_uint8, C.c_int, u32p]
    lib.tqo_uncompress_block_unsorted.restype = sz
    lib.tqo_uncompress_block_unsorted.argtypes = [u8p, C.c_uint8, C.c_int, u32p]
    lib.tqo_vint_compress_sorted.restype = sz
    lib.tqo_vint_compress_sorted.argtypes = [u32p, sz, C.c_uint32, u8p]
    lib.tqo_vint_compress_unsorted.restype = sz
    lib.tqo_vint_compress_unsorted.argtypes = [u32p, sz, u8p]
    lib.tqo_vint_uncompress_sorted.restype = sz
    lib.tqo_vint_uncompress_sorted.argtypes = [u8p, sz, C.c_uint32, C.c_uint32, u32p]
    lib.tqo_search_block.restype = sz
    lib.tqo_search_block.argtypes = [u32p, C.c_uint32]
    lib.tqo_encode_bitwidth.restype = C.c_uint8
    lib.tqo_encode_bitwidth.argtypes = [C.c_uint8, C.c_int]
    lib.tqo_encode_block_wand_max_tf.restype = C.c_uint8
    lib.tqo_encode_block_wand_max_tf.argtypes = [C.c_uint32]
    lib.tqo_decode_block_wand_max_tf.restype = C.c_uint32
    lib.tqo_decode_block_wand_max_tf.argtypes = [C.c_uint8]
    lib.tqo_bm25_idf.restype = C.c_float
    lib.tqo_bm25_idf.argtypes = [C.c_uint64, C.c_uint64]
    lib.tqo_bm25_weight.restype = C.c_float
    lib.tqo_bm25_weight.argtypes = [C.c_uint64, C.c_uint64, C.c_float]
    lib.tqo_bm25_tf_cache.argtypes = [C.c_float, f32p]
    lib.tqo_id_to_fieldnorm.restype = C.c_uint32
    lib.tqo_id_to_fieldnorm.argtypes = [C.c_uint8]
    lib.tqo_fieldnorm_to_id.restype = C.c_uint8
    lib.tqo_fieldnorm_to_id.argtypes = [C.c_uint32]
    lib.tqo_field_writer_create.restype = vp
    lib.tqo_field_writer_create.argtypes = [C.c_int, C.c_uint64, u8p, C.c_uint32]
    lib.tqo_field_writer_add_term.argtypes = [vp, u32p, u32p, C.c_uint32, u64p, u64p]
    lib.tqo_field_writer_body.argtypes = [vp, C.POINTER(u8p), C.POINTER(sz)]
    lib.tqo_field_writer_destroy.argtypes = [vp]
    lib.tqo_top_n_heap.restype = sz
    lib.tqo_top_n_heap.argtypes = [f32p, u32p, sz, sz, f32p, f32p, u32p]
    lib.tqo_merge_top_k.restype = sz
    lib.tqo_merge_top_k.argtypes = [f32p, u32p, u32p, sz, sz, sz, f32p, u32p, u32p]
    return lib


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = _load()
    return _lib


def _u32(a):
    return np.ascontiguousarray(a, dtype=np.uint32)


class FieldWriter:
    """PostingsSerializer restatement (oracle side) — builds one field's `.idx` body."""

    def __init__(self, record_option, total_num_tokens, fieldnorm_ids, max_doc):
        self._fn = None if fieldnorm_ids is None else np.ascontiguousarray(fieldnorm_ids, dtype=np.uint8)
        self.h = lib().tqo_field_writer_create(record_option, int(total_num_tokens), ptr(self._fn, u8p), max_doc)

    def add_term(self, docs, tfs=None):
        docs = _u32(docs)
        tfs = None if tfs is None else _u32(tfs)
        s, e = C.c_uint64(), C.c_uint64()
        rc = lib().tqo_field_writer_add_term(self.h, ptr(docs, u32p), ptr(tfs, u32p), len(docs), C.byref(s), C.byref(e))
        assert rc == 0
        return int(s.value), int(e.value)

    def body(self):
        p, n = u8p(), C.c_size_t()
        lib().tqo_field_writer_body(self.h, C.byref(p), C.byref(n))
        return np.ctypeslib.as_array(p, shape=(n.value,)).copy()

    def __del__(self):
        if getattr(self, "h", None):
            lib().tqo_field_writer_destroy(self.h)
            self.h = None


class OracleIndex:
    """CPU oracle with the product's segment_register / search_batch surface."""

    def __init__(self):
        self.h = lib().tqo_index_create()
        self._keep = []

    def segment_register(self, segment_ord, field, max_doc, record_option, idx_body, fieldnorm=None, alive=None):
        idx_body = np.ascontiguousarray(idx_body, dtype=np.uint8)
        fn = None if fieldnorm is None else np.ascontiguousarray(fieldnorm, dtype=np.uint8)
        al = None if alive is None else np.ascontiguousarray(alive, dtype=np.uint8)
        rc = lib().tqo_segment_register(self.h, segment_ord, field, max_doc, record_option, ptr(idx_body, u8p), idx_body.size,
                                        ptr(fn, u8p), 0 if fn is None else fn.size, ptr(al, u8p), 0 if al is None else al.size)
        assert rc == 0

    def search_batch(self, batch: QueryBatch, mode=0, n_threads=1):
        stride, scores, segs, docs, counts = batch.alloc_out()
        rc = lib().tqo_search_batch(self.h, batch.ptr, batch.nq, mode, n_threads, stride, ptr(scores, f32p), ptr(segs, u32p),
                                    ptr(docs, u32p), ptr(counts, u32p))
        if rc != 0:
            raise RuntimeError(lib().tqo_last_error(self.h).decode())
        return scores, segs, docs, counts

    def register_positions(self, segment_ord, field, pos_bytes):
        """The field's `.pos` sub-file of a registered segment (phrase queries)."""
        p = np.ascontiguousarray(pos_bytes, dtype=np.uint8)
        rc = lib().tqo_segment_register_positions(self.h, segment_ord, field, ptr(p, u8p), len(p))
        if rc != 0:
            raise RuntimeError(lib().tqo_last_error(self.h).decode())

    def phrase_search(self, terms, weight, avg_fieldnorm, slop=0, cap=4096):
        """PhraseScorer over the registered segments.  terms: (offset in phrase, segment_ord, field, doc_freq,
        postings_start, postings_end, positions_start, positions_end) per (term, segment).
        Returns [(segment, doc, score, phrase_count)] in (segment, doc) order."""
        arr = np.zeros(len(terms), dtype=PHRASE_TERM_DTYPE)
        for i, t in enumerate(terms):
            arr[i] = tuple(int(x) for x in t)
        n_phrase_terms = len({int(t[0]) for t in terms})
        sg, dc = np.zeros(cap, np.uint32), np.zeros(cap, np.uint32)
        sc, ct = np.zeros(cap, np.float32), np.zeros(cap, np.uint32)
        n = C.c_size_t(0)
        rc = lib().tqo_phrase_search(self.h, arr.ctypes.data, len(terms), n_phrase_terms, np.float32(weight), np.float32(avg_fieldnorm), slop,
                                     cap, ptr(sg, u32p), ptr(dc, u32p), ptr(sc, f32p), ptr(ct, u32p), C.byref(n))
        if rc != 0:
            raise RuntimeError(lib().tqo_last_error(self.h).decode())
        k = min(int(n.value), cap)
        return [(int(sg[i]), int(dc[i]), float(sc[i]), int(ct[i])) for i in range(k)]

    def count_batch(self, batch: QueryBatch):
        """Count collector: alive docs matching each query (src/collector/count_collector.rs)."""
        out = np.zeros(max(batch.nq, 1), dtype=np.uint64)
        rc = lib().tqo_count_batch(self.h, batch.ptr, batch.nq, ptr(out, u64p))
        if rc != 0:
            raise RuntimeError(lib().tqo_last_error(self.h).decode())
        return out[:batch.nq]

    def decode_postings(self, term_seg):
        ts = TermSeg(*[int(x) for x in term_seg])
        docs = np.zeros(max(ts.doc_freq, 1), dtype=np.uint32)
        tfs = np.zeros(max(ts.doc_freq, 1), dtype=np.uint32)
        rc = lib().tqo_decode_postings(self.h, C.byref(ts), ptr(docs, u32p), ptr(tfs, u32p))
        if rc != 0:
            raise RuntimeError(lib().tqo_last_error(self.h).decode())
        return docs[:ts.doc_freq], tfs[:ts.doc_freq]

    def block_table(self, term_seg, weight, avg_fieldnorm):
        ts = TermSeg(*[int(x) for x in term_seg])
        n = ts.doc_freq // 128
        last = np.zeros(max(n, 1), dtype=np.uint32)
        bm = np.zeros(max(n, 1), dtype=np.float32)
        rc = lib().tqo_block_table(self.h, C.byref(ts), weight, avg_fieldnorm, ptr(last, u32p), ptr(bm, f32p))
        assert rc == 0
        return last[:n], bm[:n]

    def term_scorer(self, term_seg, weight, avg_fieldnorm):
        return TermScorer(self, term_seg, weight, avg_fieldnorm)

    def __del__(self):
        if getattr(self, "h", None):
            lib().tqo_index_destroy(self.h)
            self.h = None


class TermScorer:
    """Handle on the restated TermScorer (src/query/term_query/term_scorer.rs)."""

    def __init__(self, index, term_seg, weight, avg_fieldnorm):
        self._ix = index
        ts = TermSeg(*[int(x) for x in term_seg])
        self.h = lib().tqo_term_scorer_open(index.h, C.byref(ts), weight, avg_fieldnorm)
        assert self.h

    def doc(self): return lib().tqo_term_scorer_doc(self.h)
    def advance(self): return lib().tqo_term_scorer_advance(self.h)
    def seek(self, t): return lib().tqo_term_scorer_seek(self.h, t)
    def seek_block(self, t): lib().tqo_term_scorer_seek_block(self.h, t)
    def term_freq(self): return lib().tqo_term_scorer_term_freq(self.h)
    def last_doc_in_block(self): return lib().tqo_term_scorer_last_doc_in_block(self.h)
    def score(self): return lib().tqo_term_scorer_score(self.h)
    def block_max_score(self): return lib().tqo_term_scorer_block_max_score(self.h)
    def max_score(self): return lib().tqo_term_scorer_max_score(self.h)

    def __del__(self):
        if getattr(self, "h", None):
            lib().tqo_term_scorer_close(self.h)
            self.h = None


def bm25_idf(doc_freq, doc_count): return lib().tqo_bm25_idf(doc_freq, doc_count)
def bm25_weight(doc_freq, doc_count, boost=1.0): return lib().tqo_bm25_weight(doc_freq, doc_count, boost)


def bm25_tf_cache(avg):
    out = np.zeros(256, dtype=np.float32)
    lib().tqo_bm25_tf_cache(avg, ptr(out, f32p))
    return out


def id_to_fieldnorm(i): return lib().tqo_id_to_fieldnorm(i)
def fieldnorm_to_id(f): return lib().tqo_fieldnorm_to_id(f)


def top_n_heap(scores, docs, k):
    scores = np.ascontiguousarray(scores, dtype=np.float32)
    docs = _u32(docs)
    n = len(scores)
    thr = np.zeros(max(n, 1), dtype=np.float32)
    os_, od = np.zeros(max(k, 1), dtype=np.float32), np.zeros(max(k, 1), dtype=np.uint32)
    m = lib().tqo_top_n_heap(ptr(scores, f32p), ptr(docs, u32p), n, k, ptr(thr, f32p), ptr(os_, f32p), ptr(od, u32p))
    return thr[:n], list(zip(os_[:m].tolist(), od[:m].tolist()))


def merge_top_k(hits, start, end):
    sc = np.ascontiguousarray([h[0] for h in hits], dtype=np.float32)
    sg = _u32([h[1] for h in hits])
    dc = _u32([h[2] for h in hits])
    n = len(hits)
    os_, og, od = np.zeros(max(end, 1), np.float32), np.zeros(max(end, 1), np.uint32), np.zeros(max(end, 1), np.uint32)
    m = lib().tqo_merge_top_k(ptr(sc, f32p), ptr(sg, u32p), ptr(dc, u32p), n, start, end, ptr(os_, f32p), ptr(og, u32p), ptr(od, u32p))
    return list(zip(os_[:m].tolist(), og[:m].tolist(), od[:m].tolist()))


PHRASE_TERM_DTYPE = np.dtype([("offset", "<u4"), ("segment_ord", "<u4"), ("field", "<u4"), ("doc_freq", "<u4"),
                              ("postings_start", "<u8"), ("postings_end", "<u8"), ("positions_start", "<u8"), ("positions_end", "<u8")])


# ---- positions codec (src/positions, N3 groundwork) -------------------------------------------------------------
def positions_serialize(deltas, chunk=0):
    """PositionSerializer: write_positions_delta (in chunks of `chunk`) + close_term -> the term's bytes."""
    d = np.ascontiguousarray(deltas, dtype=np.uint32)
    n = lib().tqo_positions_serialize(ptr(d, u32p), len(d), chunk, None, 0)
    out = np.zeros(max(n, 1), dtype=np.uint8)
    lib().tqo_positions_serialize(ptr(d, u32p), len(d), chunk, ptr(out, u8p), n)
    return out[:n]


class PositionReader:
    def __init__(self, data):
        self._data = np.ascontiguousarray(data, dtype=np.uint8)
        self.h = lib().tqo_position_reader_open(ptr(self._data, u8p), len(self._data))
        if not self.h:
            raise ValueError("corrupt positions data")

    def read(self, offset, n):
        out = np.zeros(max(n, 1), dtype=np.uint32)
        lib().tqo_position_reader_read(self.h, offset, ptr(out, u32p), n)
        return out[:n]

    def close(self):
        if self.h:
            lib().tqo_position_reader_close(self.h)
            self.h = None

    __del__ = close
