// The shared-decode query engine: every posting list a batch touches is decoded and scored ONCE per batch, and all
// queries of the batch are evaluated against doc-id tiles of those scored postings held in shared memory.
//
// Why: a (term, doc) BM25 score = Bm25Weight.weight * tf/(tf + norm[fieldnorm_id]) (src/query/bm25.rs:158-175) does not
// depend on the query that asks for it -- the weight belongs to the term.  A 512-query batch of 5-term unions over a
// 100M-doc index asks for 14 G scored postings, but the index only holds 0.3 G: the per-query kernels (k_or_strip ...)
// decode every list ~47 times per step.  Here:
//
//   k_score_lists   K1 + K2 of SURVEY.md §2, once per distinct (list, weight, tf-norm table) of the batch: one warp per
//                   128-doc block, BitPacker4x unpack + strict-delta prefix sum (BlockDecoder::uncompress_block_sorted /
//                   _unsorted, src/postings/compression/mod.rs:105-150), fieldnorm gather (src/fieldnorm/reader.rs:128),
//                   BM25 -> (doc, score) pairs streamed to HBM scratch with 16-byte stores.
//   k_tile          K3/K4/K5/K6 for all queries of the batch at once: a CTA walks a range of 1024-doc tiles; per tile it
//                   stages the pairs of every list into shared memory (sorted by doc inside a list), then evaluates the
//                   queries against them:
//                     * MaxScore with per-tile maxima (the block-max idea of block_wand_union.rs:16-43,171-177 at tile
//                       granularity, with maxima computed from the scores themselves, so they are exact under the
//                       searcher's global statistics -- the stored block-max pair is only exact under the segment's own
//                       average fieldnorm, term_scorer.rs:58-70): clauses are kept in the canonical summation order
//                       (descending weight); the longest suffix whose tile maxima add up to less than the query's
//                       threshold is non-essential and is only LOOKED UP for docs that the essential clauses make
//                       promising;
//                     * the essential postings of all (query, tile) pairs become one work list that the CTA's threads share
//                       evenly: one thread = one posting = one doc, completed by lookups in the query's other clauses (a bit
//                       test + popcount for dense lists, a short binary search for sparse ones), the non-essential suffix
//                       with an early exit as soon as the rest of the bound cannot reach the threshold; pairs with very
//                       many essential postings (and the sample launch) take a warp-private window of f32 score slots
//                       instead (BufferedUnionScorer's shape, union/buffered_union.rs:63-86);
//                     * f32 sums are taken clause by clause in the canonical order, so pruned and exhaustive evaluation
//                       give bit-identical scores;
//                     * survivors (score key >= the query's threshold) go to the query's candidate region, k_final
//                       (TopNHeap / merge_fruits) selects exactly.
//   thresholds      a first launch over 1/16 of the tiles only SAMPLES scores (the complete scores of the docs that hold one of
//                   a union's two heaviest clauses / a conjunction's matches); the k-th best sample is a valid lower bound of
//                   the final k-th score (k_theta_samples).  The exact launches follow,
//                   each ending in k_theta (exact k-th best candidate so far), so most tiles are visited under a
//                   near-final threshold.  Sampled tiles are visited again by the exact launches: nothing is pushed twice.
//
// Everything is bit-exact w.r.t. the oracle's exhaustive mode: same per-posting IEEE operations, same summation order.
#pragma once
#include "tq_kernels.cuh"

namespace tq {

constexpr uint32_t kTile = 1024;            // docs per tile
constexpr uint32_t kTileThreads = 256;
constexpr uint32_t kTileWarps = kTileThreads / 32;
constexpr uint32_t kTileMaxQueries = 1024;  // (query, segment) pairs of one group in one segment
constexpr uint32_t kTileMaxSlots = 4096;    // distinct scored lists of one group in one segment
constexpr uint32_t kTileMaxBig = 64;        // dense lists of one segment that get a tile index and a presence map
constexpr uint32_t kTileMaxPairs = 12288;   // pairs of one tile held in shared memory (start | len are 16-bit fields)
constexpr uint32_t kTileExactWindows = 2;   // warps that take heavy pairs in an exact launch (each owns a window of kTile f32 slots)
constexpr uint32_t kSampleSeg = 6;          // sample launch: postings scored per driving clause and (query, tile)
constexpr uint32_t kSamplePerTile = 4;      // sample launch: the best few partial maxima of a (query, tile)
constexpr uint32_t kNoSlot = 0xFFFFFFFFu;
constexpr uint32_t kTileOpAnd = 1;  // TQ_OP_AND; term queries and unions share one evaluation (a union of one clause)
constexpr uint32_t kTileOpBool = 3;  // TQ_OP_BOOL: the clause words are [n_groups, need_should, n_should, n_not, (len, slots..) per MUST group,
                                     // SHOULD slots.., MUST_NOT slots..], groups by ascending cost, slots inside by descending weight

struct TSlot {  // one distinct scored list of a group: (posting list, Bm25Weight.weight, tf-norm table)
  uint32_t list_id;
  float weight;
  uint32_t cache_idx;
  uint32_t doc_freq;
  uint32_t pair_base;  // first element of this list's pairs in p_docs / p_scores (multiple of 128)
  uint32_t big;        // column in the segment's tile index, kNoSlot for small lists
  uint32_t tseg;
  uint32_t pad;
};
struct TSeg {  // one segment of a group
  uint32_t slot_base, n_slots, n_big;  // slots [0, n_big) have a tile index and are staged by whole warps
  uint32_t query_base, n_queries;
  uint32_t max_doc, segment_ord, n_tiles;
  uint32_t clause_base, n_clause_words;  // this segment's share of TileParams::clauses (staged in shared memory when it fits)
  uint32_t pad0, pad1;
  const uint8_t* alive;
  uint32_t* tix;  // [(n_tiles + 1)][n_big]: index of the list's first pair with doc >= tile * kTile (written by k_score_lists)
};
struct TQuery {  // one (query, segment): what Collector::collect_segment sees
  uint32_t query;
  uint32_t clause_base;  // into TileParams::clauses: slot ordinals (local to the segment) in canonical summation order
  uint16_t n_clauses;
  uint8_t op;
  uint8_t flags;  // bit 0: prunable (every weight finite and >= 0); bit 1: phrase (op = AND; `query` is the QSeg ordinal of the phrase)
};
struct TUnit { uint32_t tseg, t0, t1, pad; };
struct SChunk { uint32_t slot, b0, b1; };  // a CTA's share of k_score_lists: blocks [b0, b1) of one slot (global slot ordinal)

struct TileParams {
  const TSlot* slots;
  const TSeg* segs;
  const TQuery* queries;
  const uint16_t* clauses;
  const TUnit* units;
  const SChunk* chunks;
  uint32_t* p_docs;
  float* p_scores;
  uint32_t* samples;       // [n_queries of the batch][sample_cap] score keys
  uint32_t* sample_count;  // [n_queries of the batch]
  uint32_t* flags;         // [0]: a tile held more pairs than p_cap (the batch is re-run on the per-query kernels)
  unsigned long long* counters;  // [8] diagnostics: 0 (query, tile) pairs seen, 1 skipped (bound / no essential posting / cold window),
                                 //     2 light pairs, 3 heavy pairs, 4 essential postings applied, 5 docs completed, 6 docs at or above the threshold
  uint32_t sample_cap;
  uint32_t p_cap;      // pairs of one tile that fit in shared memory
  uint32_t max_slots;  // largest n_slots of any segment of the group
  uint32_t max_big;    // ... n_big
  uint32_t max_queries;  // ... n_queries
  uint32_t cl_cap;     // clause words staged in shared memory per CTA (0: the segment's clauses are read from global memory)
  uint32_t seg_cap;    // entries of the per-tile work list (essential clauses of the flat pairs)
  uint32_t light_max;  // (query, tile) pairs with at most this many essential postings are evaluated posting by posting (flat path);
                       // above it (and with more than one essential clause) a warp accumulates the query's window
  uint32_t n_win;      // warps of a CTA that own a window (2: heavy pairs are rare; 8 for groups with wide queries)
  // phrase queries on the tile engine: conjunctions whose matches are handed to k_phrase_verify as records
  //   [qseg, doc, (global slot, pair index) per term]  (pair index: into p_docs / p_scores / p_pos)
  uint32_t has_phrase; // the group holds phrase queries (k_tile keeps every slot's first staged pair index on chip)
  uint32_t ph_stride;  // words per record = 2 + 2 * (most terms of a phrase in the batch)
  uint32_t ph_cap;     // records the buffer holds (more: flags[2], the batch is repeated on the per-query kernels)
  uint32_t pad_ph;
  uint32_t* p_pos;     // per pair of a slot with TSlot::pad & 1: sum of the term frequencies before it in its 128-doc block
  uint32_t* ph_recs;
  uint32_t* ph_count;
};


// ---- K1 + K2, once per batch -------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads, 4) k_score_lists(const BatchParams P, const TileParams TP, uint32_t chunk_base) {
  const SChunk C = TP.chunks[chunk_base + blockIdx.x];
  const TSlot sl = TP.slots[C.slot];
  const ListDesc L = P.lists[sl.list_id];
  const Scorer sc{sl.weight, P.caches + 256u * sl.cache_idx, P.tf_tables + (size_t)(kTfRows * 256u) * sl.cache_idx};
  const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  uint32_t* out_docs = TP.p_docs + sl.pair_base;
  float* out_scores = TP.p_scores + sl.pair_base;
  uint32_t* tix = nullptr;
  uint32_t n_big = 0, n_tiles = 0;
  const bool want_pos = TP.p_pos != nullptr && (sl.pad & 1u);
  if (sl.big != kNoSlot) {
    const TSeg G = TP.segs[sl.tseg];
    tix = G.tix + sl.big;
    n_big = G.n_big;
    n_tiles = G.n_tiles;
  }
  BlockFetch f;
  uint32_t b = C.b0 + warp;
  if (b < C.b1) fetch_issue(L, b, lane, f);
  for (; b < C.b1; b += kWarps) {
    uint32_t doc[4], tf[4];
    fetch_decode(L, b, f, lane, doc, tf);
    const uint32_t prev_last = b ? __ldg(L.last_doc + b - 1) : 0xFFFFFFFFu;
    if (b + kWarps < C.b1) fetch_issue(L, b + kWarps, lane, f);  // the next block travels while this one is scored
    const uint32_t g0 = b * 128u + lane * 4u;
    uint32_t id[4];
    float s[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const bool valid = g0 + i < L.doc_freq;  // the VInt tail pads with TERMINATED
      id[i] = L.fieldnorm ? (uint32_t)__ldg(L.fieldnorm + (valid ? doc[i] : 0u)) : 1u;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const bool valid = g0 + i < L.doc_freq;
      s[i] = valid ? bm25_score_id(sc, id[i], tf[i]) : 0.0f;
      if (!valid) doc[i] = 0xFFFFFFFFu;
    }
    // the pairs are read back once, by another kernel, after the whole pass: stream them past the L2 lines the gathers re-use
    // (fieldnorm bytes, block tables)
    __stcs(reinterpret_cast<uint4*>(out_docs + g0), make_uint4(doc[0], doc[1], doc[2], doc[3]));
    __stcs(reinterpret_cast<float4*>(out_scores + g0), make_float4(s[0], s[1], s[2], s[3]));
    if (want_pos) {  // where the doc's positions start inside the block's share of the position stream (segment_postings.rs:232-254)
      uint32_t tfv[4], pre[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) tfv[i] = g0 + i < L.doc_freq ? tf[i] : 0u;
      const uint32_t sum4 = tfv[0] + tfv[1] + tfv[2] + tfv[3];
      const uint32_t incl = warp_incl_scan(sum4, lane);
      pre[0] = incl - sum4; pre[1] = pre[0] + tfv[0]; pre[2] = pre[1] + tfv[1]; pre[3] = pre[2] + tfv[2];
      __stcs(reinterpret_cast<uint4*>(TP.p_pos + sl.pair_base + g0), make_uint4(pre[0], pre[1], pre[2], pre[3]));
    }
    if (tix) {  // tile index of a dense list: tix[t] = first pair with doc >= t * kTile
      uint32_t pd = __shfl_up_sync(kFull, doc[3], 1);
      if (lane == 0) pd = prev_last;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const uint32_t g = g0 + i;
        if (g < L.doc_freq) {
          const uint32_t p = i == 0 ? pd : doc[i - 1];
          const uint32_t t_cur = doc[i] / kTile;
          for (uint32_t t = p == 0xFFFFFFFFu ? 0u : p / kTile + 1u; t <= t_cur; ++t) tix[(size_t)t * n_big] = g;
          if (g + 1u == L.doc_freq)
            for (uint32_t t = t_cur + 1u; t <= n_tiles; ++t) tix[(size_t)t * n_big] = L.doc_freq;
        }
      }
    }
  }
}

// ---- tile evaluation -----------------------------------------------------------------------------------------------------
// What a tile keeps in shared memory per slot: info = start | len << 16 (pairs [start, start + len) of s_off / s_score, ascending
// docs), the largest score, and for DENSE slots (the first n_big of a segment) a 1024-bit presence map with per-word ranks, so that
// "does this list hold doc x, and where" is one bit test + one popcount instead of a binary search over hundreds of pairs.
struct TileView {
  const uint16_t* s_off;
  const float* s_score;
  const uint32_t* s_info;
  const uint32_t* s_bits;   // [max_big][32]
  const uint16_t* s_rank;   // [max_big][32] pairs of the slot before word w
  const uint32_t* s_mask;   // [max_slots] sparse slots: bit j set = the slot lists a doc of the tile's j-th 32-doc stripe
  uint32_t n_big;
};

__device__ __forceinline__ float tile_find(const TileView& V, uint32_t slot, uint32_t off, bool& found) {
  const uint32_t info = V.s_info[slot];
  if (slot < V.n_big) {
    const uint32_t w = V.s_bits[slot * 32u + (off >> 5)];
    const uint32_t bit = 1u << (off & 31u);
    found = (w & bit) != 0u;
    if (!found) return 0.0f;
    return V.s_score[(info & 0xFFFFu) + V.s_rank[slot * 32u + (off >> 5)] + __popc(w & (bit - 1u))];
  }
  // sparse slot: most lookups miss -- one stripe bit answers them; the rest scans the few pairs of the slot
  found = false;
  if (!((V.s_mask[slot] >> (off >> 5)) & 1u)) return 0.0f;
  uint32_t lo = info & 0xFFFFu;
  const uint32_t end = lo + (info >> 16);
  if (end - lo > 8u) {
    uint32_t hi = end;
    while (hi - lo > 4u) {
      const uint32_t m = (lo + hi) >> 1;
      if (V.s_off[m] < off) lo = m + 1u; else hi = m;
    }
  }
  for (; lo < end; ++lo) {
    const uint32_t o = V.s_off[lo];
    if (o >= off) { found = o == off; break; }
  }
  return found ? V.s_score[lo] : 0.0f;
}

__device__ __forceinline__ void tile_push(const BatchParams& P, uint32_t query, float score, uint32_t doc, uint32_t segment_ord,
                                          uint32_t th_key, const uint8_t* __restrict__ alive) {
  const uint32_t key = score_to_key(score);
  if (key < th_key) return;
  if (alive && !is_alive(alive, doc)) return;
  const DQuery Q = P.queries[query];
  const uint32_t idx = atomicAdd(&P.qstate[query].cand_count, 1u);
  if (idx < Q.cand_cap) P.cands[Q.cand_base + idx] = Cand{key, segment_ord, doc, 0u};
}

// One (query, tile) pair that is evaluated posting by posting: what every posting's thread needs to know (16 bytes).
struct TileQ {
  uint16_t qi;           // the query: index into the CTA's copy of the segment's queries (batch ordinal, clause slots)
  uint8_t n_op;          // clauses [0:6) | op << 6
  uint8_t ne_prune;      // essential prefix (unions) / driving clause (conjunctions) [0:6) | phrase << 6 | prune << 7
  uint32_t th_key;       // the query's threshold when the tile was entered
  float ne;              // bound of the non-essential suffix in this tile (unions); 0 for conjunctions
  uint32_t shared_stripes;  // unions: 32-doc stripes of the tile in which more than one essential clause lists a doc (a posting
                            // outside them is the only essential posting of its doc: no lookup is needed to know its partial sum)
  __device__ __forceinline__ uint32_t n() const { return n_op & 63u; }
  __device__ __forceinline__ uint32_t op() const { return n_op >> 6; }
  __device__ __forceinline__ uint32_t n_e() const { return ne_prune & 63u; }
  __device__ __forceinline__ bool prune() const { return (ne_prune & 128u) != 0; }
};
// the segment's queries as the CTA keeps them on chip (8 bytes)
struct TileTQ {
  uint32_t query;
  uint16_t clause_base;  // local to the segment
  uint8_t n_clauses;
  uint8_t op_flags;      // op [0:2) | phrase << 6 | prunable << 7
};
struct TileSeg { uint16_t q, clause, start, len; };  // postings [start, start + len) of one essential clause of flat pair q

// Union: the doc at tile offset `off`, met as posting p of essential clause c.  Exact score in the canonical order, or nothing when
// the bounds show it cannot reach the threshold.  Returns true when `sum` is a complete score.
__device__ __forceinline__ bool tile_eval_or(const TileView& V, const float* __restrict__ s_max, const uint16_t* __restrict__ cl, const TileQ& q,
                                             uint32_t c, uint32_t p, uint32_t off, float theta_f, float& sum) {
  bool f;
  const uint32_t n_e = q.n_e(), n = q.n();
  const bool prune = q.prune();
  sum = V.s_score[p];  // (-0.0 + s == s: SumCombiner starts from 0, score_combiner.rs:39-57)
  if ((q.shared_stripes >> (off >> 5)) & 1u) {  // another essential clause lists a doc near this one: look the doc up
    for (uint32_t c1 = 0; c1 < c; ++c1) {  // an earlier essential clause lists the doc: it is evaluated there
      tile_find(V, cl[c1], off, f);
      if (f) return false;
    }
    for (uint32_t c2 = c + 1u; c2 < n_e; ++c2) {
      const float v = tile_find(V, cl[c2], off, f);
      if (f) sum = __fadd_rn(sum, v);
    }
  }
  // the non-essential suffix, densest last; stop as soon as what is left cannot lift the doc over the threshold
  float rem = q.ne;
  for (uint32_t c2 = n_e; c2 < n; ++c2) {
    if (prune && (sum + rem) * 1.00001f < theta_f) return false;
    const uint32_t slot = cl[c2];
    rem -= s_max[slot];
    const float v = tile_find(V, slot, off, f);
    if (f) sum = __fadd_rn(sum, v);
  }
  return true;
}

// Conjunction: the doc (posting p of the driving clause) must be listed by every clause; the score is summed in clause order
// (leader = rarest list first, block_wand_intersection.rs:27,146-158) whichever clause drives the iteration.
__device__ __forceinline__ bool tile_eval_and(const TileView& V, const uint16_t* __restrict__ cl, uint32_t n, uint32_t drv, uint32_t p, uint32_t off,
                                              float& sum) {
  sum = 0.0f;
  for (uint32_t c = 0; c < n; ++c) {
    float v;
    if (c == drv) v = V.s_score[p];
    else {
      bool f;
      v = tile_find(V, cl[c], off, f);
      if (!f) return false;
    }
    sum = c == 0 ? v : __fadd_rn(sum, v);
  }
  return true;
}

// Mixed boolean query (BooleanWeight::complex_scorer for term leaves, boolean_weight.rs:236-431): the doc at `off` is met as posting p
// of the clause at word `wi`, one of the driving words [d0, d0 + dn).  It is evaluated once (by the first driving clause that lists
// it): every MUST group needs a clause that lists it (group score = sum of its matching clauses), MUST_NOT clauses must not, at least
// `need` SHOULD clauses must; score = (sum of the group scores, ascending cost) + (sum of the matching SHOULD clauses)
// (Intersection::score, RequiredOptionalScorer::score).
__device__ __forceinline__ bool tile_eval_bool(const TileView& V, const uint16_t* __restrict__ cl, uint32_t d0, uint32_t wi, uint32_t p, uint32_t off,
                                               float& sum) {
  bool f;
  for (uint32_t i = d0; i < wi; ++i) {
    tile_find(V, cl[i], off, f);
    if (f) return false;
  }
  const uint32_t ng = cl[0], need = cl[1], ns = cl[2], nn = cl[3];
  uint32_t w = 4;
  float total = 0.0f;
  for (uint32_t g = 0; g < ng; ++g) {
    const uint32_t glen = cl[w++];
    float gs = 0.0f;
    bool any = false;
    for (uint32_t e = 0; e < glen; ++e, ++w) {
      float v;
      if (w == wi) { v = V.s_score[p]; f = true; } else v = tile_find(V, cl[w], off, f);
      if (f) { gs = any ? __fadd_rn(gs, v) : v; any = true; }
    }
    if (!any) return false;
    total = g ? __fadd_rn(total, gs) : gs;
  }
  float ss = 0.0f;
  uint32_t cnt = 0;
  for (uint32_t e = 0; e < ns; ++e, ++w) {
    float v;
    if (w == wi) { v = V.s_score[p]; f = true; } else v = tile_find(V, cl[w], off, f);
    if (f) { ss = cnt ? __fadd_rn(ss, v) : v; ++cnt; }
  }
  for (uint32_t e = 0; e < nn; ++e, ++w) {
    tile_find(V, cl[w], off, f);
    if (f) return false;
  }
  if (cnt < need) return false;
  sum = ng ? (cnt ? __fadd_rn(total, ss) : total) : ss;
  return true;
}

constexpr uint32_t kTileStageWords = 2u * kTileMaxBig + kTileWarps * 17u;
// Shared memory of a k_tile CTA: the work list (TileQ records + segments) and two windows for the rare heavy pairs ...
__host__ __device__ constexpr size_t tile_union_bytes(uint32_t max_queries, uint32_t seg_cap, uint32_t n_win) {
  // (the segment list doubles as stage A's scratch: never smaller than kTileStageWords words)
  const size_t seg_bytes = (size_t)seg_cap * sizeof(TileSeg) > kTileStageWords * 4u ? (size_t)seg_cap * sizeof(TileSeg) : kTileStageWords * 4u;
  return ((size_t)n_win * kTile * 4 + (size_t)max_queries * sizeof(TileQ) + seg_bytes + 15) & ~(size_t)15;
}
__host__ __device__ constexpr size_t tile_smem_bytes(uint32_t p_cap, uint32_t max_slots, uint32_t max_big, uint32_t max_queries, uint32_t seg_cap,
                                                      uint32_t cl_cap, uint32_t n_win, uint32_t has_phrase = 0) {
  return (size_t)p_cap * 4 + tile_union_bytes(max_queries, seg_cap, n_win) + (size_t)max_slots * 20 + (size_t)max_big * 128 +
         (size_t)max_queries * sizeof(TileTQ) + (size_t)max_big * 64 + (size_t)p_cap * 2 + (size_t)max_queries * 2 + (size_t)cl_cap * 2 + 64 +
         (has_phrase ? (size_t)max_slots * 4 + 8 : 0);
}

__global__ void __launch_bounds__(kTileThreads, 3) k_tile(const BatchParams P, const TileParams TP, uint32_t unit_base, uint32_t sample_mode) {
  extern __shared__ __align__(16) unsigned char s_dyn[];
  float* s_score = reinterpret_cast<float*>(s_dyn);                            // [p_cap]
  unsigned char* s_union = reinterpret_cast<unsigned char*>(s_score + TP.p_cap);
  const size_t union_bytes = tile_union_bytes(TP.max_queries, TP.seg_cap, TP.n_win);
  const uint32_t n_win = TP.n_win;
  float* s_acc = reinterpret_cast<float*>(s_union);                            // [n_win][kTile] windows of the heavy pairs
  TileQ* s_q = reinterpret_cast<TileQ*>(s_acc + n_win * kTile);                // [max_queries] flat pairs of this tile
  TileSeg* s_seg = reinterpret_cast<TileSeg*>(s_q + TP.max_queries);           // [seg_cap]
  uint32_t* s_info = reinterpret_cast<uint32_t*>(s_union + union_bytes);       // [max_slots]
  float* s_max = reinterpret_cast<float*>(s_info + TP.max_slots);              // [max_slots] largest score of the slot in this tile (>= 0)
  uint32_t* s_cur = reinterpret_cast<uint32_t*>(s_max + TP.max_slots);         // [max_slots] small slots: first pair not staged yet
  uint32_t* s_nxt = s_cur + TP.max_slots;                                      // [max_slots] ... and its doc (0xFFFFFFFF: list exhausted)
  uint32_t* s_mask = s_nxt + TP.max_slots;                                     // [max_slots] stripe masks of the sparse slots
  uint32_t* s_bits = s_mask + TP.max_slots;                                    // [max_big][32]
  TileTQ* s_tq = reinterpret_cast<TileTQ*>(s_bits + TP.max_big * 32u);         // [max_queries] this segment's queries
  uint16_t* s_rank = reinterpret_cast<uint16_t*>(s_tq + TP.max_queries);       // [max_big][32]
  uint16_t* s_off = s_rank + TP.max_big * 32u;                                 // [p_cap]
  uint16_t* s_heavy = s_off + TP.p_cap;                                        // [max_queries]
  uint16_t* s_cl = s_heavy + TP.max_queries;                                   // [cl_cap] this segment's clause slots
  // [max_slots] phrase groups only: index (into the pair arrays) of the slot's first pair staged for the current tile
  uint32_t* s_ord0 = TP.has_phrase ? reinterpret_cast<uint32_t*>((reinterpret_cast<uintptr_t>(s_cl + TP.cl_cap) + 3u) & ~(uintptr_t)3u) : nullptr;
  __shared__ uint32_t s_total, s_nheavy, s_hpos, s_nflat, s_nseg, s_segpos, s_segvalid, s_ndefer;
  __shared__ uint32_t s_defer[kTileMaxQueries / 32u];  // queries that did not fit this round's work list
  __shared__ unsigned long long s_stat[8];
  // stage A scratch, in the (then idle) segment list: the dense slots' ranges of this tile and each warp's prefix over its slots
  uint32_t* s_ba = reinterpret_cast<uint32_t*>(s_seg);                         // [kTileMaxBig] first pair of the slot in this tile
  uint32_t* s_bn = s_ba + kTileMaxBig;                                         // [kTileMaxBig] ... and how many
  uint32_t (*s_wpre)[9] = reinterpret_cast<uint32_t (*)[9]>(s_bn + kTileMaxBig);
  uint32_t (*s_wa)[8] = reinterpret_cast<uint32_t (*)[8]>(s_bn + kTileMaxBig + kTileWarps * 9u);
  static_assert(2u * kTileMaxBig + kTileWarps * 17u <= kTileStageWords, "stage A scratch");
  const TUnit U = TP.units[unit_base + blockIdx.x];
  const TSeg G = TP.segs[U.tseg];
  const TSlot* __restrict__ slots = TP.slots + G.slot_base;
  const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
  const float neg_zero = __uint_as_float(0x80000000u);
  const TileView V{s_off, s_score, s_info, s_bits, s_rank, s_mask, G.n_big};
  for (uint32_t i = tid; i < n_win * kTile; i += kTileThreads) s_acc[i] = neg_zero;
  if (tid < 8) s_stat[tid] = 0ull;
  if (tid < kTileMaxQueries / 32u) s_defer[tid] = 0u;
  // the segment's queries and their clause slots are read for every tile: keep them on chip (clause_base becomes an index into cl0)
  const bool cl_staged = G.n_clause_words <= TP.cl_cap;
  const uint16_t* __restrict__ cl0 = cl_staged ? s_cl : TP.clauses + G.clause_base;
  for (uint32_t i = tid; i < G.n_queries; i += kTileThreads) {
    const TQuery tq = TP.queries[G.query_base + i];
    s_tq[i] = TileTQ{tq.query, (uint16_t)(tq.clause_base - G.clause_base), (uint8_t)tq.n_clauses, (uint8_t)((tq.op & 3u) | ((tq.flags & 2u) << 5) | ((tq.flags & 1u) << 7))};
  }
  if (cl_staged)
    for (uint32_t i = tid; i < G.n_clause_words; i += kTileThreads) s_cl[i] = TP.clauses[G.clause_base + i];
  // small slots: position the cursor on the first pair at or after the unit's first doc.  Four slots per thread at a time: the
  // binary searches advance in lockstep, so their (dependent) loads overlap -- with two-tile units this prologue is a tenth of the CTA's time
  for (uint32_t s0 = G.n_big + tid; s0 < G.n_slots; s0 += 4u * kTileThreads) {
    const uint32_t lo0 = U.t0 * kTile;
    const uint32_t* d[4];
    uint32_t a[4], b[4], df[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const uint32_t s = s0 + (uint32_t)j * kTileThreads;
      const bool in = s < G.n_slots;
      const TSlot sl = slots[in ? s : s0];
      d[j] = TP.p_docs + sl.pair_base;
      df[j] = sl.doc_freq;
      a[j] = 0; b[j] = in ? sl.doc_freq : 0u;
    }
    for (;;) {
      bool any = false;
      uint32_t v[4], m[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        m[j] = (a[j] + b[j]) >> 1;
        if (a[j] < b[j]) { v[j] = __ldg(d[j] + m[j]); any = true; }
      }
      if (!any) break;
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (a[j] < b[j]) { if (v[j] < lo0) a[j] = m[j] + 1u; else b[j] = m[j]; }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const uint32_t s = s0 + (uint32_t)j * kTileThreads;
      if (s < G.n_slots) {
        s_cur[s] = a[j];
        s_nxt[s] = a[j] < df[j] ? __ldg(d[j] + a[j]) : 0xFFFFFFFFu;
      }
    }
  }
  __syncthreads();
  unsigned long long st_pairs = 0, st_skip = 0, st_flat = 0, st_heavy = 0, st_ess = 0, st_compl = 0, st_push = 0;  // per-thread diagnostics

  for (uint32_t t = U.t0; t < U.t1; ++t) {
    const uint32_t lo = t * kTile, hi = lo + kTile;
    if (tid < G.n_big) {  // three independent loads per dense slot: one round trip for the whole tile
      const uint32_t a = __ldg(G.tix + (size_t)t * G.n_big + tid), b = __ldg(G.tix + (size_t)(t + 1u) * G.n_big + tid);
      s_ba[tid] = __ldg(&slots[tid].pair_base) + a;
      s_bn[tid] = b - a;
    }
    if (tid == 0) { s_total = 0; s_nheavy = 0; s_hpos = 0; s_nflat = 0; s_nseg = 0; s_segpos = 0; s_segvalid = TP.seg_cap; s_ndefer = 0; }
    __syncthreads();
    // ---- stage A: this tile's pairs of every slot -> shared memory --------------------------------------------------------
    if (warp < G.n_big) {
      // dense lists: warp w owns slots w, w+8, ... (<= 8 of them).  Their [tix[t], tix[t+1]) ranges are laid end to end and copied as
      // ONE flat loop, so the loads of all the warp's slots are in flight together (one DRAM round trip per 128 pairs instead
      // of three dependent ones per slot)
      const uint32_t ne = (G.n_big - warp + kTileWarps - 1u) / kTileWarps;
      uint32_t n_e = 0;
      if (lane < ne) {
        const uint32_t s = warp + lane * kTileWarps;
        n_e = s_bn[s];
        s_wa[warp][lane] = s_ba[s];
      }
      const uint32_t incl = warp_incl_scan(n_e, lane);
      if (lane < 8u) s_wpre[warp][lane + 1u] = incl;
      if (lane == 0) s_wpre[warp][0] = 0u;
      const uint32_t total_w = __shfl_sync(kFull, incl, 7);
      uint32_t base = 0;
      if (lane == 0 && total_w) base = atomicAdd(&s_total, total_w);
      base = __shfl_sync(kFull, base, 0);
      for (uint32_t e = 0; e < ne; ++e) s_bits[(warp + e * kTileWarps) * 32u + lane] = 0u;
      __syncwarp();
      const bool fits = base + total_w <= TP.p_cap;
      if (fits) {
        for (uint32_t i0 = 0; i0 < total_w; i0 += 128u) {
          uint32_t dd[4], ss[4];
          float vv[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const uint32_t i = i0 + (uint32_t)u * 32u + lane;
            dd[u] = 0xFFFFFFFFu;
            if (i < total_w) {
              uint32_t e = 0;
              while (i >= s_wpre[warp][e + 1u]) ++e;
              ss[u] = warp + e * kTileWarps;
              const size_t g = (size_t)s_wa[warp][e] + (i - s_wpre[warp][e]);
              dd[u] = __ldg(TP.p_docs + g);
              vv[u] = __ldg(TP.p_scores + g);
            }
          }
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const uint32_t i = i0 + (uint32_t)u * 32u + lane;
            if (dd[u] != 0xFFFFFFFFu) {
              const uint32_t off = dd[u] - lo;
              s_off[base + i] = (uint16_t)off;
              s_score[base + i] = vv[u];
              atomicOr(&s_bits[ss[u] * 32u + (off >> 5)], 1u << (off & 31u));
            }
          }
        }
      }
      __syncwarp();
      for (uint32_t e = 0; e < ne; ++e) {
        const uint32_t s = warp + e * kTileWarps;
        const uint32_t b = base + s_wpre[warp][e], n = s_wpre[warp][e + 1u] - s_wpre[warp][e];
        const uint32_t cnt = (uint32_t)__popc(s_bits[s * 32u + lane]);
        s_rank[s * 32u + lane] = (uint16_t)(warp_incl_scan(cnt, lane) - cnt);
        float mx = 0.0f;
        if (fits)
          for (uint32_t i = lane; i < n; i += 32) mx = fmaxf(mx, s_score[b + i]);
        mx = __uint_as_float(__reduce_max_sync(kFull, __float_as_uint(mx)));  // mx >= 0
        if (lane == 0) {
          s_info[s] = (b & 0xFFFFu) | (n << 16); s_max[s] = mx;
          if (s_ord0) s_ord0[s] = s_wa[warp][e];
        }
      }
    }
    for (uint32_t s = G.n_big + tid; s < G.n_slots; s += kTileThreads) {  // sparse lists: one thread walks its cursor
      uint32_t nd = s_nxt[s];
      if (nd >= hi) {
        s_info[s] = 0u; s_max[s] = 0.0f; s_mask[s] = 0u;
        continue;
      }
      const TSlot sl = slots[s];
      const uint32_t* __restrict__ d = TP.p_docs + sl.pair_base;
      const float* __restrict__ sc = TP.p_scores + sl.pair_base;
      const uint32_t cur = s_cur[s];
      // the pair at the cursor belongs to this tile (nd < hi); fetch it and the next three, docs and scores, in ONE round trip
      uint32_t d4[4];
      float v4[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const bool in = cur + i < sl.doc_freq;
        d4[i] = in ? __ldg(d + cur + i) : 0xFFFFFFFFu;
        v4[i] = in ? __ldg(sc + cur + i) : 0.0f;
      }
      uint32_t n = 0;
#pragma unroll
      for (int i = 0; i < 4; ++i) n += d4[i] < hi ? 1u : 0u;  // ascending: the ones below hi are a prefix
      if (n == 4u) {  // (rare for a sparse list) keep walking
        for (;;) {
          const uint32_t c = cur + n;
          uint32_t e4[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) e4[i] = c + i < sl.doc_freq ? __ldg(d + c + i) : 0xFFFFFFFFu;
          uint32_t k = 0;
#pragma unroll
          for (int i = 0; i < 4; ++i) k += e4[i] < hi ? 1u : 0u;
          n += k;
          if (k < 4u) { nd = e4[k]; break; }
        }
      } else {
        nd = d4[n];
      }
      s_cur[s] = cur + n;
      s_nxt[s] = nd;
      if (nd < hi + kTile) {  // the next tile reads this list again: start its pairs on their way to L2 now
        asm volatile("prefetch.global.L2 [%0];" ::"l"(d + cur + n));
        asm volatile("prefetch.global.L2 [%0];" ::"l"(sc + cur + n));
      }
      const uint32_t base = atomicAdd(&s_total, n);
      float mx = 0.0f;
      uint32_t mask = 0;
      if (base + n <= TP.p_cap) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          if ((uint32_t)i < n) {
            const uint32_t off = d4[i] - lo;
            s_off[base + i] = (uint16_t)off;
            s_score[base + i] = v4[i];
            mask |= 1u << (off >> 5);
            mx = fmaxf(mx, v4[i]);
          }
        }
        for (uint32_t i = 4; i < n; ++i) {
          const float v = __ldg(sc + cur + i);
          const uint32_t off = __ldg(d + cur + i) - lo;
          s_off[base + i] = (uint16_t)off;
          s_score[base + i] = v;
          mask |= 1u << (off >> 5);
          mx = fmaxf(mx, v);
        }
      }
      s_info[s] = (base & 0xFFFFu) | (n << 16);
      s_max[s] = mx;
      s_mask[s] = mask;
      if (s_ord0) s_ord0[s] = sl.pair_base + cur;
    }
    __syncthreads();
    if (s_total > TP.p_cap) {  // more pairs than the tile buffer holds: give the batch back to the per-query kernels
      if (tid == 0) atomicExch(TP.flags, 1u);
      __syncthreads();
      continue;
    }
    // ---- stage B1: one thread per query -- bounds and routing ---------------------------------------------------------------
    // skipped: no doc of the tile can reach the query's threshold;  flat: the essential postings become work items that the CTA's
    // threads share evenly (B2);  heavy: a warp accumulates the query's window (B3; sample launches, and pairs with many postings).
    // (a work list that fills up defers the remaining pairs to another round of B1 + B2: the list is sized for the common case)
    for (uint32_t round = 0;; ++round) {
    for (uint32_t qi = tid; qi < G.n_queries; qi += kTileThreads) {
      if (round) {
        const uint32_t bit = 1u << (qi & 31u);
        if (!(s_defer[qi >> 5] & bit)) continue;
        atomicAnd(&s_defer[qi >> 5], ~bit);
      }
      const TileTQ tq = s_tq[qi];
      const uint16_t* __restrict__ cl = cl0 + tq.clause_base;
      const uint32_t th_key = *(volatile unsigned int*)&P.qstate[tq.query].theta;
      const float theta_f = threshold_score(th_key);
      const bool prune = (tq.op_flags & 128u) && theta_f > 0.0f && !sample_mode;
      const uint32_t n = tq.n_clauses, op = tq.op_flags & 3u;
      ++st_pairs;
      TileQ rec;
      rec.qi = (uint16_t)qi; rec.th_key = th_key; rec.n_op = (uint8_t)((op == kTileOpBool ? 0u : n) | (op << 6));
      uint32_t cnt = 0, n_segs = 0, rec_ne = 0;
      uint32_t b_d0 = 0, b_dn = 0;  // kTileOpBool: the driving words
      if (op == kTileOpBool) {
        const uint32_t ng = cl[0], ns = cl[2];
        uint32_t w = 4, best = 0xFFFFFFFFu;
        float bound = 0.0f;
        bool dead = false;
        for (uint32_t g = 0; g < ng; ++g) {  // the MUST group with the fewest postings in this tile drives
          const uint32_t glen = cl[w];
          uint32_t tot = 0;
          for (uint32_t e = 0; e < glen; ++e) { const uint32_t slot = cl[w + 1u + e]; tot += s_info[slot] >> 16; bound += s_max[slot]; }
          if (tot == 0u) dead = true;
          if (tot < best) { best = tot; b_d0 = w + 1u; b_dn = glen; }
          w += 1u + glen;
        }
        uint32_t s_tot = 0;
        for (uint32_t e = 0; e < ns; ++e) { const uint32_t slot = cl[w + e]; s_tot += s_info[slot] >> 16; bound += s_max[slot]; }
        if (ng == 0u) { best = s_tot; b_d0 = w; b_dn = ns; }
        if (dead || best == 0u || (prune && bound * 1.00001f < theta_f)) { ++st_skip; continue; }
        cnt = best;
        for (uint32_t e = 0; e < b_dn; ++e) n_segs += (s_info[cl[b_d0 + e]] >> 16) ? 1u : 0u;
        rec.ne = __uint_as_float(b_d0 | (b_dn << 16));
        rec.shared_stripes = 0u;
      } else if (op == kTileOpAnd) {
        float bound = 0.0f;
        uint32_t dmin = 0xFFFFFFFFu, drv = 0;
        for (uint32_t c = 0; c < n; ++c) {  // the clause with the fewest postings in this tile drives
          const uint32_t len = s_info[cl[c]] >> 16;
          bound += s_max[cl[c]];
          if (len < dmin) { dmin = len; drv = c; }
        }
        if (dmin == 0u || (prune && bound * 1.00001f < theta_f)) { ++st_skip; continue; }
        rec_ne = drv; rec.ne = 0.0f; rec.shared_stripes = 0u;
        cnt = dmin; n_segs = 1;
      } else {
        // sample launch: only the docs of the two heaviest clauses are scored (completely): the best docs of a union nearly
        // always hold its rarest terms, and ANY real score is a valid sample
        uint32_t n_e = n;
        if (sample_mode) {  // ... extended by further clauses while the tile holds fewer than four of their postings
          uint32_t have = 0;
          n_e = 0;
          while (n_e < n && (n_e < 2u || have < 4u)) { have += s_info[cl[n_e]] >> 16; ++n_e; }
        }
        float ne = 0.0f;
        if (prune) {
          while (n_e > 0) {
            const float b = ne + s_max[cl[n_e - 1u]];
            if (!(b * 1.00001f < theta_f)) break;  // (f32 sums of <= 32 terms differ by < 4e-6 relative between orders)
            ne = b;
            --n_e;
          }
        }
        if (n_e == 0) { ++st_skip; continue; }  // no doc of this tile can reach the threshold
        uint32_t seen = 0, shared = 0;
        for (uint32_t c = 0; c < n_e; ++c) {
          const uint32_t slot = cl[c];
          const uint32_t len = s_info[slot] >> 16;
          cnt += len; n_segs += len ? 1u : 0u;
          const uint32_t m = slot < G.n_big ? (len ? 0xFFFFFFFFu : 0u) : s_mask[slot];  // (dense lists: every stripe)
          shared |= seen & m;
          seen |= m;
        }
        if (cnt == 0) { ++st_skip; continue; }  // a doc without an essential posting stays below the threshold
        rec_ne = n_e; rec.ne = ne; rec.shared_stripes = shared;
      }
      rec.ne_prune = (uint8_t)(rec_ne | (prune ? 128u : 0u));
      // sample launch: ANY set of real scores is a valid sample -- at most kSampleSeg postings per driving clause, no window path
      if (sample_mode && *(volatile unsigned int*)&TP.sample_count[tq.query] >= TP.sample_cap) continue;
      bool heavy = !sample_mode && cnt > TP.light_max && op != kTileOpAnd && op != kTileOpBool && rec_ne >= 2u;
      if (n_segs > TP.seg_cap) heavy = true;  // (test hook sizes only)
      if (!heavy) {
        const uint32_t sb = atomicAdd(&s_nseg, n_segs);
        if (sb + n_segs > TP.seg_cap) {  // the work list is full (entries from here on are not written): next round
          atomicMin(&s_segvalid, sb);
          atomicOr(&s_defer[qi >> 5], 1u << (qi & 31u));
          s_ndefer = 1u;
        } else {
          const uint32_t qslot = atomicAdd(&s_nflat, 1u);
          s_q[qslot] = rec;
          uint32_t w = sb;
          const uint32_t len_cap = sample_mode ? kSampleSeg : 0xFFFFu;
          if (op == kTileOpBool) {
            for (uint32_t e = 0; e < b_dn; ++e) {
              const uint32_t info = s_info[cl[b_d0 + e]];
              if (info >> 16) s_seg[w++] = TileSeg{(uint16_t)qslot, (uint16_t)(b_d0 + e), (uint16_t)(info & 0xFFFFu), (uint16_t)min(info >> 16, len_cap)};
            }
          } else if (op == kTileOpAnd) {
            const uint32_t info = s_info[cl[rec_ne]];
            s_seg[w] = TileSeg{(uint16_t)qslot, (uint16_t)rec_ne, (uint16_t)(info & 0xFFFFu), (uint16_t)min(info >> 16, len_cap)};
          } else {
            for (uint32_t c = 0; c < rec_ne; ++c) {
              const uint32_t info = s_info[cl[c]];
              if (info >> 16) s_seg[w++] = TileSeg{(uint16_t)qslot, (uint16_t)c, (uint16_t)(info & 0xFFFFu), (uint16_t)min(info >> 16, len_cap)};
            }
          }
          ++st_flat;
          st_ess += cnt;
        }
      }
      if (heavy) s_heavy[atomicAdd(&s_nheavy, 1u)] = (uint16_t)qi;
    }
    __syncthreads();
    const bool more_rounds = s_ndefer != 0u;
    // ---- stage B2: the flat pairs' essential postings, one per thread --------------------------------------------------------
    {
      const uint32_t n_seg = min(s_nseg, s_segvalid);
      for (;;) {
        uint32_t sbase = 0;
        if (lane == 0) sbase = atomicAdd(&s_segpos, 32u);
        sbase = __shfl_sync(kFull, sbase, 0);
        if (sbase >= n_seg) break;
        // 32 segments, one per lane; their postings are dealt to the lanes 32 at a time (load-balanced expansion)
        TileSeg mine = TileSeg{0, 0, 0, 0};
        if (sbase + lane < n_seg) mine = s_seg[sbase + lane];
        const uint32_t incl = warp_incl_scan((uint32_t)mine.len, lane);
        const uint32_t total = __shfl_sync(kFull, incl, 31);
        const uint32_t excl = incl - mine.len;
        for (uint32_t ibase = 0; ibase < total; ibase += 32u) {
          const uint32_t item = ibase + lane;
          // the segment of item `ibase + lane`: every listed segment holds a posting, so segment numbers grow by one at each
          // start; (segments that start before this chunk) - 1 + (starts at or before the lane's item)
          const uint32_t rel = excl - ibase;  // (wraps for segments that start before the chunk)
          const uint32_t starts = __reduce_or_sync(kFull, (mine.len && rel < 32u) ? (1u << rel) : 0u);
          const uint32_t earlier = (uint32_t)__popc(__ballot_sync(kFull, mine.len && excl < ibase));
          const uint32_t j = earlier - 1u + (uint32_t)__popc(starts & (0xFFFFFFFFu >> (31u - lane)));
          const uint32_t incl_j = __shfl_sync(kFull, incl, j & 31u);
          uint32_t s_query = 0xFFFFFFFFu - lane, s_key = 0;  // sample launch: this lane's sample (if any), handed over below
          if (item < total) {
            const TileSeg sg = s_seg[sbase + j];
            const uint32_t before = incl_j - sg.len;
            const uint32_t p = sg.start + (item - before);
            const TileQ q = s_q[sg.q];
            const TileTQ qq = s_tq[q.qi];
            const uint16_t* __restrict__ cl = cl0 + qq.clause_base;
            const uint32_t off = s_off[p];
            float sum;
            bool ok;
            if (q.op() == kTileOpBool) ok = tile_eval_bool(V, cl, __float_as_uint(q.ne) & 0xFFFFu, sg.clause, p, off, sum);
            else if (q.op() == kTileOpAnd) ok = tile_eval_and(V, cl, q.n(), q.n_e(), p, off, sum);
            else ok = tile_eval_or(V, s_max, cl, q, sg.clause, p, off, threshold_score(q.th_key), sum);
            if (ok) {
              ++st_compl;
              const uint32_t key = score_to_key(sum);
              if (key >= q.th_key) ++st_push;
              if (!sample_mode) tile_push(P, qq.query, sum, lo + off, G.segment_ord, q.th_key, G.alive);
              else if (key >= q.th_key && key != 0u && (!G.alive || is_alive(G.alive, lo + off))) { s_query = qq.query; s_key = key; }  // (a deleted doc bounds nothing)
            }
          }
          if (sample_mode) {  // one atomic per query and warp step instead of one per sample (the counters are hot)
            const unsigned peers = __match_any_sync(kFull, s_query);
            if (s_key) {
              const uint32_t leader = (uint32_t)__ffs(peers) - 1u;
              uint32_t base = 0;
              if (lane == leader) base = atomicAdd(&TP.sample_count[s_query], (uint32_t)__popc(peers));
              base = __shfl_sync(peers, base, leader);
              const uint32_t idx = base + (uint32_t)__popc(peers & lanemask_lt(lane));
              if (idx < TP.sample_cap) TP.samples[(size_t)s_query * TP.sample_cap + idx] = s_key;
            }
          }
        }
      }
    }
    if (!more_rounds) break;
    __syncthreads();
    if (tid == 0) { s_nflat = 0; s_nseg = 0; s_segpos = 0; s_segvalid = TP.seg_cap; s_ndefer = 0; }
    __syncthreads();
    }
    // ---- stage B3: one warp per heavy pair, a window of f32 score slots ----------------------------------------------------
    {
      const uint32_t n_heavy = s_nheavy;
      float* acc = s_acc + warp * kTile;
      for (; warp < n_win;) {  // (the warps without a window go straight to the barrier)
        uint32_t h = 0;
        if (lane == 0) h = atomicAdd(&s_hpos, 1u);
        h = __shfl_sync(kFull, h, 0);
        if (h >= n_heavy) break;
        const TileTQ tq = s_tq[s_heavy[h]];
        const uint16_t* __restrict__ cl = cl0 + tq.clause_base;
        uint32_t th_key = 0;
        if (lane == 0) th_key = *(volatile unsigned int*)&P.qstate[tq.query].theta;
        th_key = __shfl_sync(kFull, th_key, 0);  // one value for the whole warp (the query-wide threshold moves)
        const float theta_f = threshold_score(th_key);
        const bool prune = (tq.op_flags & 128u) && theta_f > 0.0f && !sample_mode;
        const uint32_t n = tq.n_clauses;
        if (lane == 0) ++st_heavy;
        if ((tq.op_flags & 3u) == kTileOpBool) {  // lanes share the driving group's postings (as for conjunctions below)
          const uint32_t ng = cl[0], ns = cl[2];
          uint32_t w = 4, best_tot = 0xFFFFFFFFu, d0 = 0, dn = 0;
          for (uint32_t g = 0; g < ng; ++g) {
            const uint32_t glen = cl[w];
            uint32_t tot = 0;
            for (uint32_t e = 0; e < glen; ++e) tot += s_info[cl[w + 1u + e]] >> 16;
            if (tot < best_tot) { best_tot = tot; d0 = w + 1u; dn = glen; }
            w += 1u + glen;
          }
          if (ng == 0u) { d0 = w; dn = ns; }
          uint32_t best = 0;
          for (uint32_t e = 0; e < dn; ++e) {
            const uint32_t info = s_info[cl[d0 + e]];
            const uint32_t a = info & 0xFFFFu, en = a + (info >> 16);
            for (uint32_t p = a + lane; p < en; p += 32) {
              const uint32_t off = s_off[p];
              float sum;
              if (!tile_eval_bool(V, cl, d0, d0 + e, p, off, sum)) continue;
              if (sample_mode) { if (!G.alive || is_alive(G.alive, lo + off)) best = max(best, score_to_key(sum)); }
              else { ++st_compl; tile_push(P, tq.query, sum, lo + off, G.segment_ord, th_key, G.alive); }
            }
            if (lane == 0) st_ess += en - a;
          }
          if (sample_mode) {
            for (uint32_t r = 0; r < kSamplePerTile; ++r) {
              const uint32_t mk = __reduce_max_sync(kFull, best);
              if (mk == 0u || mk < th_key) break;
              const unsigned who = __ballot_sync(kFull, best == mk);
              if (lane == (uint32_t)__ffs(who) - 1u) {
                const uint32_t idx = atomicAdd(&TP.sample_count[tq.query], 1u);
                if (idx < TP.sample_cap) TP.samples[(size_t)tq.query * TP.sample_cap + idx] = mk;
                best = 0;
              }
            }
          }
          __syncwarp();
          continue;
        }
        if ((tq.op_flags & 3u) == kTileOpAnd) {  // lanes share the driving clause's postings; every lane looks its docs up in the others
          uint32_t dmin = 0xFFFFFFFFu, drv = 0;
          for (uint32_t c = 0; c < n; ++c) {
            const uint32_t len = s_info[cl[c]] >> 16;
            if (len < dmin) { dmin = len; drv = c; }
          }
          const uint32_t info = s_info[cl[drv]];
          const uint32_t a = info & 0xFFFFu, e = a + (info >> 16);
          uint32_t best = 0;
          for (uint32_t p = a + lane; p < e; p += 32) {
            const uint32_t off = s_off[p];
            float sum;
            if (!tile_eval_and(V, cl, n, drv, p, off, sum)) continue;
            if (sample_mode) { if (!G.alive || is_alive(G.alive, lo + off)) best = max(best, score_to_key(sum)); }  // (a deleted doc bounds nothing)
            else { ++st_compl; tile_push(P, tq.query, sum, lo + off, G.segment_ord, th_key, G.alive); }
          }
          if (lane == 0) st_ess += e - a;
          if (sample_mode) {
            for (uint32_t r = 0; r < kSamplePerTile; ++r) {
              const uint32_t mk = __reduce_max_sync(kFull, best);
              if (mk == 0u || mk < th_key) break;
              const unsigned who = __ballot_sync(kFull, best == mk);
              if (lane == (uint32_t)__ffs(who) - 1u) {
                const uint32_t idx = atomicAdd(&TP.sample_count[tq.query], 1u);
                if (idx < TP.sample_cap) TP.samples[(size_t)tq.query * TP.sample_cap + idx] = mk;
                best = 0;
              }
            }
          }
          __syncwarp();
          continue;
        }
        uint32_t n_e = n;
        float ne = 0.0f;
        if (prune) {
          while (n_e > 0) {
            const float b = ne + s_max[cl[n_e - 1u]];
            if (!(b * 1.00001f < theta_f)) break;
            ne = b;
            --n_e;
          }
        }
        float wmax = 0.0f;
        for (uint32_t c = 0; c < n_e; ++c) {  // clause order is the f32 summation order
          const uint32_t info = s_info[cl[c]];
          const uint32_t a = info & 0xFFFFu, e = a + (info >> 16);
          for (uint32_t p = a + lane; p < e; p += 32) {
            const uint32_t o = s_off[p];
            const float v = __fadd_rn(acc[o], s_score[p]);
            acc[o] = v;
            wmax = fmaxf(wmax, v);
          }
          if (lane == 0) st_ess += e - a;
          __syncwarp();
        }
        if (sample_mode) {
          // the best partial maximum of every lane's slots; the top few of the warp are this (query, tile)'s samples
          uint32_t best = 0;
          for (uint32_t g = 0; g < kTile / 128u; ++g) {
            const uint32_t idx = g * 128u + lane * 4u;
            const float4 v = *reinterpret_cast<const float4*>(acc + idx);
            *reinterpret_cast<float4*>(acc + idx) = make_float4(neg_zero, neg_zero, neg_zero, neg_zero);
            const float vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int c = 0; c < 4; ++c)
              if (__float_as_uint(vv[c]) != 0x80000000u && (!G.alive || is_alive(G.alive, lo + idx + c))) best = max(best, score_to_key(vv[c]));  // a deleted doc's score bounds nothing
          }
          for (uint32_t r = 0; r < kSamplePerTile; ++r) {
            const uint32_t m = __reduce_max_sync(kFull, best);
            if (m == 0u || m < th_key) break;
            const unsigned who = __ballot_sync(kFull, best == m);
            if (lane == (uint32_t)__ffs(who) - 1u) {
              const uint32_t idx = atomicAdd(&TP.sample_count[tq.query], 1u);
              if (idx < TP.sample_cap) TP.samples[(size_t)tq.query * TP.sample_cap + idx] = m;
              best = 0;
            }
          }
          __syncwarp();
          continue;
        }
        // nothing of this window can reach the threshold when even its largest partial sum plus the bound cannot
        const float mx = __uint_as_float(__reduce_max_sync(kFull, __float_as_uint(wmax)));
        const bool cold = prune && (mx + ne) * 1.00001f < theta_f;
        if (cold && lane == 0) ++st_skip;
        for (uint32_t g = 0; g < kTile / 128u; ++g) {
          const uint32_t idx = g * 128u + lane * 4u;
          float4 v = make_float4(neg_zero, neg_zero, neg_zero, neg_zero);
          if (!cold) v = *reinterpret_cast<const float4*>(acc + idx);
          *reinterpret_cast<float4*>(acc + idx) = make_float4(neg_zero, neg_zero, neg_zero, neg_zero);
          if (cold) continue;
          const float vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            float sum = vv[c];
            if (__float_as_uint(sum) == 0x80000000u) continue;  // untouched
            const uint32_t off = idx + c;
            float rem = ne;
            bool dead = false;
            for (uint32_t c2 = n_e; c2 < n; ++c2) {  // the non-essential suffix, with the same early exit as the flat path
              if (prune && (sum + rem) * 1.00001f < theta_f) { dead = true; break; }
              const uint32_t slot = cl[c2];
              rem -= s_max[slot];
              bool f;
              const float x = tile_find(V, slot, off, f);
              if (f) sum = __fadd_rn(sum, x);
            }
            if (dead) continue;
            ++st_compl;
            if (score_to_key(sum) >= th_key) ++st_push;
            tile_push(P, tq.query, sum, lo + off, G.segment_ord, th_key, G.alive);
          }
        }
        __syncwarp();
      }
    }
    __syncthreads();
  }
  if (TP.counters) {
    // warp-aggregated diagnostics
    unsigned long long v[7] = {st_pairs, st_skip, st_flat, st_heavy, st_ess, st_compl, st_push};
#pragma unroll
    for (int i = 0; i < 7; ++i) {
      unsigned long long x = v[i];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(kFull, x, o);
      if (lane == 0 && x) atomicAdd(&s_stat[i], x);
    }
    __syncthreads();
    if (tid < 7 && s_stat[tid]) atomicAdd(&TP.counters[tid], s_stat[tid]);
  }
}

// ---- thresholds from score keys -------------------------------------------------------------------------------------------
// k-th largest of n u32 keys (4-pass radix select), block-wide; returns 0 when n < k.  `need_out`: how many keys equal to the
// result are needed on top of the strictly larger ones to make k.
__device__ __forceinline__ uint32_t block_kth_key(const uint32_t* __restrict__ keys, uint32_t n, uint32_t stride_words, uint32_t k,
                                                  uint32_t* s_hist, uint32_t* s_prefix, uint32_t* s_need) {
  if (n < k) return 0u;
  if (threadIdx.x == 0) { *s_prefix = 0; *s_need = k; }
  uint32_t mask = 0;
  for (int pass = 0; pass < 4; ++pass) {
    const int shift = 24 - 8 * pass;
    for (uint32_t i = threadIdx.x; i < 256; i += blockDim.x) s_hist[i] = 0;
    __syncthreads();
    const uint32_t prefix = *s_prefix;
    for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) {
      const uint32_t key = keys[(size_t)i * stride_words];
      if ((key & mask) == prefix) atomicAdd(&s_hist[(key >> shift) & 255u], 1u);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      uint32_t need = *s_need, bsel = 0;
      for (int bin = 255; bin >= 0; --bin) {
        const uint32_t h = s_hist[bin];
        if (h >= need) { bsel = (uint32_t)bin; break; }
        need -= h;
      }
      *s_need = need;
      *s_prefix = prefix | (bsel << shift);
    }
    mask |= 0xFFu << shift;
    __syncthreads();
  }
  return *s_prefix;
}

// After the sample launch: the k-th best sampled score of a query bounds its final k-th score from below.
__global__ void __launch_bounds__(kThreads) k_theta_samples(const BatchParams P, const uint32_t* __restrict__ samples,
                                                            const uint32_t* __restrict__ sample_count, uint32_t sample_cap) {
  __shared__ uint32_t s_hist[256];
  __shared__ uint32_t s_prefix, s_need;
  const uint32_t q = blockIdx.x;
  const uint32_t n = min(sample_count[q], sample_cap);
  const uint32_t kth = block_kth_key(samples + (size_t)q * sample_cap, n, 1u, P.queries[q].k, s_hist, &s_prefix, &s_need);
  if (threadIdx.x == 0 && kth) atomicMax(&P.qstate[q].theta, kth);
}

// Cross-shard exchange, export side: the k best score keys a query holds so far (samples after the sample launch, else
// candidates), k_stride entries per query, zero padded.  The caller all-gathers these over the shards.
__global__ void __launch_bounds__(kThreads) k_topkeys_export(const BatchParams P, const uint32_t* __restrict__ samples,
                                                             const uint32_t* __restrict__ sample_count, uint32_t sample_cap,
                                                             uint32_t from_samples, uint32_t* __restrict__ out, uint32_t k_stride) {
  __shared__ uint32_t s_hist[256];
  __shared__ uint32_t s_prefix, s_need, s_n;
  const uint32_t q = blockIdx.x;
  const DQuery Q = P.queries[q];
  const uint32_t* keys;
  uint32_t n, stride;
  if (from_samples && samples) { keys = samples + (size_t)q * sample_cap; n = min(sample_count[q], sample_cap); stride = 1u; }
  else { keys = &P.cands[Q.cand_base].score_key; n = min(P.qstate[q].cand_count, Q.cand_cap); stride = sizeof(Cand) / 4u; }
  const uint32_t k = min(Q.k, k_stride);
  uint32_t* o = out + (size_t)q * k_stride;
  for (uint32_t i = threadIdx.x; i < k_stride; i += blockDim.x) o[i] = 0u;
  if (threadIdx.x == 0) s_n = 0;
  __syncthreads();
  if (Q.op == 3u) return;  // TQ_OP_PHRASE: its candidates may still be arriving (second stream); no threshold exchange for phrases
  if (n <= k) {
    for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) o[i] = keys[(size_t)i * stride];
    return;
  }
  const uint32_t kth = block_kth_key(keys, n, stride, k, s_hist, &s_prefix, &s_need);
  const uint32_t above = k - s_need;  // strictly larger keys
  for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) {
    const uint32_t key = keys[(size_t)i * stride];
    if (key > kth) o[atomicAdd(&s_n, 1u)] = key;
  }
  __syncthreads();
  for (uint32_t i = above + threadIdx.x; i < k; i += blockDim.x) o[i] = kth;
}

// ... import side: gathered [n_shards][nq][k_stride] keys -> the exact k-th best of the union becomes the query's threshold.
__global__ void __launch_bounds__(kThreads) k_theta_from_keys(const BatchParams P, const uint32_t* __restrict__ gathered, uint32_t n_shards,
                                                              uint32_t nq, uint32_t k_stride) {
  extern __shared__ uint32_t s_keys[];  // [n_shards * k_stride]
  __shared__ uint32_t s_hist[256];
  __shared__ uint32_t s_prefix, s_need;
  const uint32_t q = blockIdx.x;
  const uint32_t n = n_shards * k_stride;
  for (uint32_t i = threadIdx.x; i < n; i += blockDim.x)
    s_keys[i] = gathered[((size_t)(i / k_stride) * nq + q) * k_stride + (i % k_stride)];
  __syncthreads();
  const uint32_t kth = block_kth_key(s_keys, n, 1u, P.queries[q].k, s_hist, &s_prefix, &s_need);
  if (threadIdx.x == 0 && kth) atomicMax(&P.qstate[q].theta, kth);
}

}  // namespace tq
