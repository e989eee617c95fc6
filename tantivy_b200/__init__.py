"""tantivy_b200 — B200-native execution of tantivy's posting decode -> AND/OR -> BM25 -> top-k path.

Importing this package loads the CUDA library (tantivy_b200/_lib/libtantivy_b200.so) and fails
loudly if it has not been built; there is no CPU fallback."""
from ._abi import (TQ_OCCUR_MUST, TQ_OCCUR_MUST_NOT, TQ_OCCUR_SHOULD, TQ_OP_BOOL, TERMINATED, TQ_MAX_K, TQ_MAX_TERMS, TQ_OP_AND, TQ_OP_OR, TQ_OP_PHRASE, TQ_OP_TERM, TQ_RECORD_BASIC, TQ_RECORD_FREQS,
                   TQ_RECORD_FREQS_POSITIONS, QueryBatch)
from .lib import (Batch, Context, FieldWriter, MultiContext, SynthIndex, TqError, bm25_idf, bm25_tf_cache, bm25_weight, fieldnorm_to_id,
                  id_to_fieldnorm)
