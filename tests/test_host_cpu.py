"""CPU-side checks of the product's host code against the oracle: the segment writer produces
byte-identical postings, the synthetic generator emits valid tantivy posting lists, BM25 scalars
agree bit for bit, and the C-ABI library exports every symbol include/tantivy_b200.h declares."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import tantivy_b200 as T
from oracle import tq_oracle as O
from tantivy_b200._abi import TQ_RECORD_BASIC, TQ_RECORD_FREQS, TQ_RECORD_FREQS_POSITIONS
from tests.helpers import OracleSegment

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_abi_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "tantivy_b200.h")).read()
    names = set(re.findall(r"\b(tq_[a-z0-9_]+)\s*\(", hdr))
    assert len(names) >= 20
    lib = C.CDLL(T.lib.SO_PATH)
    for n in sorted(names):
        assert hasattr(lib, n), f"{n} declared in the header but not exported"


def test_ctypes_structs_match_the_c_header(tmp_path):
    """sizeof / offsetof of the C ABI's structs, as a C compiler sees the header, equal the ctypes mirror's
    (tantivy_b200/_abi.py): a drifted field would silently corrupt every query."""
    import subprocess
    from tantivy_b200 import _abi as A
    src = tmp_path / "abi.c"
    fields = {"tq_term_seg": A.TermSeg, "tq_query": A.Query, "tq_stats": A.Stats}
    lines = ['#include <stdio.h>', '#include <stddef.h>', f'#include "{os.path.join(ROOT, "include", "tantivy_b200.h")}"', "int main(void) {"]
    for cname, ct in fields.items():
        lines.append(f'  printf("{cname} %zu\\n", sizeof({cname}));')
        for fname, _ in ct._fields_:
            lines.append(f'  printf("{cname}.{fname} %zu\\n", offsetof({cname}, {fname}));')
    lines += ["  return 0;", "}"]
    src.write_text("\n".join(lines))
    exe = tmp_path / "abi"
    subprocess.run(["gcc", "-std=c11", "-o", str(exe), str(src)], check=True)
    out = dict(ln.split() for ln in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.splitlines())
    for cname, ct in fields.items():
        assert int(out[cname]) == C.sizeof(ct), cname
        for fname, _ in ct._fields_:
            assert int(out[f"{cname}.{fname}"]) == getattr(ct, fname).offset, f"{cname}.{fname}"
    assert A.QUERY_DTYPE.itemsize == C.sizeof(A.Query) and A.TERM_SEG_DTYPE.itemsize == C.sizeof(A.TermSeg)
    for name in A.QUERY_DTYPE.names:
        assert A.QUERY_DTYPE.fields[name][1] == getattr(A.Query, name).offset


def test_ctx_create_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(T.TqError):
        T.Context(0)


def test_bm25_scalars_bit_identical():
    rng = np.random.default_rng(0)
    for _ in range(200):
        N = int(rng.integers(1, 10**9))
        n = int(rng.integers(1, N + 1))
        assert T.bm25_idf(n, N) == O.bm25_idf(n, N)
        assert T.bm25_weight(n, N, 1.0) == O.bm25_weight(n, N, 1.0)
        assert T.bm25_weight(n, N, 2.5) == O.bm25_weight(n, N, 2.5)
    for avg in (0.5, 2.4, 13 / 3, 95.7, 1000.0):
        assert (T.bm25_tf_cache(np.float32(avg)) == O.bm25_tf_cache(np.float32(avg))).all()
    assert [T.id_to_fieldnorm(i) for i in range(256)] == [O.id_to_fieldnorm(i) for i in range(256)]
    for f in list(range(0, 5000)) + [10**6, 2**31, 2**32 - 1]:
        assert T.fieldnorm_to_id(f) == O.fieldnorm_to_id(f)


def _random_list(rng, max_doc, n, max_gap_bits):
    if n == 0:
        return np.zeros(0, np.uint32), np.zeros(0, np.uint32)
    if max_gap_bits is None:
        docs = np.sort(rng.choice(max_doc, size=n, replace=False)).astype(np.uint32)
    else:
        gaps = rng.integers(1, 2 ** max_gap_bits + 1, size=n, dtype=np.uint64)
        docs = (np.cumsum(gaps) - 1).astype(np.uint64)
        docs = docs[docs < max_doc].astype(np.uint32)
    tfs = rng.integers(1, 12, size=len(docs)).astype(np.uint32)
    return docs, tfs


@pytest.mark.parametrize("record_option", [TQ_RECORD_BASIC, TQ_RECORD_FREQS, TQ_RECORD_FREQS_POSITIONS])
def test_writer_bytes_match_oracle(record_option):
    rng = np.random.default_rng(11 + record_option)
    max_doc = 300_000
    fieldnorms = rng.integers(1, 500, size=max_doc).astype(np.uint32)
    lists = []
    for n, bits in [(0, None), (1, None), (127, None), (128, None), (129, None), (1000, 1), (5000, 3), (4096, 5), (777, 8), (20000, None),
                    (256, 10), (300, 0)]:
        docs, tfs = _random_list(rng, max_doc, n, bits)
        if record_option == TQ_RECORD_BASIC:
            tfs = None
        lists.append((docs, tfs))
    # a block that starts at doc 0 and a list with huge tfs (32-bit tf width)
    lists.append((np.arange(0, 256, dtype=np.uint32), None if record_option == TQ_RECORD_BASIC else np.full(256, 1, np.uint32)))
    if record_option != TQ_RECORD_BASIC:
        big = np.full(130, 1, np.uint32)
        big[5] = 0xFFFFFFFF
        big[77] = 300
        lists.append((np.arange(10, 140, dtype=np.uint32), big))
    a = OracleSegment(lists, fieldnorms, record_option=record_option, writer_cls=O.FieldWriter)
    b = OracleSegment(lists, fieldnorms, record_option=record_option, writer_cls=T.FieldWriter)
    assert a.terms == b.terms
    assert a.body.tobytes() == b.body.tobytes()
    # and the oracle decodes what the product wrote
    ix = O.OracleIndex()
    b.register(ix)
    for t, (docs, tfs) in enumerate(lists):
        if len(docs) == 0:
            continue
        d, f = ix.decode_postings(b.term_seg(t))
        assert (d == docs).all()
        assert (f == (tfs if tfs is not None else 1)).all()


def test_writer_bytes_match_oracle_wide_doc_widths():
    """Doc bit widths 25..31: the two independent encoders (oracle / product writer) agree byte for byte."""
    rng = np.random.default_rng(77)
    max_doc = 0x7FFFFFFE
    lists = []
    for bits in range(25, 32):
        n = 128 + int(rng.integers(0, 100))
        gaps = rng.integers(1, 40, size=n, dtype=np.uint64)
        gaps[int(rng.integers(0, 128))] = 2 ** (bits - 1) + 1
        docs = (np.cumsum(gaps) - gaps[0]).astype(np.uint32)
        lists.append((docs, rng.integers(1, 9, size=n).astype(np.uint32)))
    a = OracleSegment(lists, None, writer_cls=O.FieldWriter, max_doc=max_doc)
    b = OracleSegment(lists, None, writer_cls=T.FieldWriter, max_doc=max_doc)
    assert a.terms == b.terms and a.body.tobytes() == b.body.tobytes()
    ix = O.OracleIndex()
    b.register(ix)
    for t, (docs, tfs) in enumerate(lists):
        d, f = ix.decode_postings(b.term_seg(t))
        assert (d == docs).all() and (f == tfs).all()


def test_writer_rejects_bad_input():
    w = T.FieldWriter(TQ_RECORD_FREQS, 10, np.ones(10, np.uint8), 10)
    with pytest.raises(T.TqError):
        w.add_term([3, 3], [1, 1])
    with pytest.raises(T.TqError):
        w.add_term([1, 2], [1, 0])
    with pytest.raises(T.TqError):
        w.add_term([1, 10], [1, 1])  # doc id >= max_doc (would index past the fieldnorm array)


def test_synth_segments_are_valid_postings():
    dens = [0.3, 0.05, 0.001, 0.00002]
    ix = T.SynthIndex(2, 150_000, dens, seed=1234, n_threads=4)
    ix2 = T.SynthIndex(2, 150_000, dens, seed=1234, n_threads=1)
    oi = O.OracleIndex()
    ix.register(oi)
    for s in range(2):
        assert ix.body(s).tobytes() == ix2.body(s).tobytes()  # independent of the thread count
        fn = ix.fieldnorm(s)
        assert len(fn) == 150_000 and fn.min() >= 1
        for t, p in enumerate(dens):
            df, st, en = ix.term_info[s][t]
            assert abs(df - p * 150_000) <= 6 * np.sqrt(p * 150_000) + 3
            if df == 0:
                continue
            docs, tfs = oi.decode_postings((0, s, 0, df, st, en))
            assert (np.diff(docs.astype(np.int64)) > 0).all() and docs[-1] < 150_000
            assert tfs.min() >= 1 and tfs.max() <= 10
            lens = np.array([O.id_to_fieldnorm(int(i)) for i in fn[docs[:200]]])
            assert (tfs[:200] <= np.maximum(lens, 1) * 2).all()
