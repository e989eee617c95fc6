// ORACLE — TEST INFRASTRUCTURE ONLY (see codec.hpp header).
// C API over the CPU restatement: the checker for tests/, smoke() and the CPU-baseline legs of
// bench.py.  Mirrors include/tantivy_b200.h (same structs) so that parity tests are one switch.
//
// Search orchestration follows (relative to /root/reference):
//   src/core/searcher.rs:220-237                     one Weight, per-segment collect, merge_fruits
//   src/collector/sort_key/sort_by_score.rs:35-66    collect_segment_top_k (TopNHeap + for_each_pruning,
//                                                    deletes checked inside the callback)
//   src/query/boolean_query/boolean_weight.rs:581-600  dispatch to block_wand / block_wand_intersection
//   src/query/term_query/term_weight.rs:118-141,179-219 TermWeight::for_each_pruning / specialized_scorer
#include <atomic>
#include <functional>
#include <map>
#include <thread>

#include "../include/tantivy_b200.h"
#include "positions.hpp"
#include "query.hpp"
#include "phrase.hpp"

using namespace tqo;

struct OSegment {
  uint32_t segment_ord, field, max_doc;
  IndexRecordOption record_option;
  std::vector<uint8_t> idx_body;  // incl. 8-byte total_num_tokens header
  std::vector<uint8_t> fieldnorm; bool has_fieldnorm;
  std::vector<uint8_t> alive; bool has_alive;
  std::vector<uint8_t> positions;  // the field's `.pos` sub-file (phrase queries, N3)
  bool is_alive(uint32_t doc) const { return !has_alive || ((alive[doc >> 3] >> (doc & 7)) & 1); }
};
struct tqo_index {
  std::map<std::pair<uint32_t, uint32_t>, OSegment> segs;
  std::string err;
};

static TermScorer make_term_scorer(const OSegment& seg, const tq_term_seg& ts, const Bm25Weight& w, uint32_t clause, bool basic_requested = false) {
  TermScorer t;
  const uint8_t* base = seg.idx_body.data() + 8 + ts.postings_start;
  const size_t len = (size_t)(ts.postings_end - ts.postings_start);
  // read_postings_from_terminfo(term_info, option).downgrade(record_option): request freqs if available
  IndexRecordOption requested = (seg.record_option == Basic || basic_requested) ? Basic : WithFreqs;
  t.postings.block_cursor = BlockSegmentPostings::open(ts.doc_freq, base, len, seg.record_option, requested);
  t.postings.cur = 0;
  t.fieldnorm_reader = seg.has_fieldnorm ? FieldNormReader::from_data(seg.fieldnorm.data(), seg.max_doc)
                                         : FieldNormReader::constant(seg.max_doc, 1);  // term_weight.rs:218
  t.similarity_weight = w;
  t.clause = clause;
  return t;
}

static Bm25Weight weight_for(const tq_query& q, uint32_t term) {
  Bm25Weight w;
  w.weight = q.weight[term];
  w.average_fieldnorm = q.avg_fieldnorm ? q.avg_fieldnorm[term] : 0.0f;
  if (q.tf_cache) std::memcpy(w.cache, q.tf_cache + 256 * (size_t)term, 256 * sizeof(float));
  else for (int id = 0; id < 256; ++id) w.cache[id] = cached_tf_component(id_to_fieldnorm((uint8_t)id), w.average_fieldnorm);
  return w;
}


// ---- BooleanQuery with mixed Occurs over term leaves ------------------------------------------------
// Restates, for TermQuery leaves and one level of all-SHOULD sub-queries under MUST, what BooleanWeight::complex_scorer builds
// (src/query/boolean_query/boolean_weight.rs:236-431) and how its scorers add up:
//   * EmptyScorers (clauses without postings in the segment) are removed first (:247-263); a MUST that is empty empties the query;
//   * minimum_number_should_match m against the n remaining SHOULD scorers (:269-301): m > n -> nothing; m = 0 -> Optional;
//     m = 1 -> Required(union); m = n (>= 2) -> the SHOULD clauses become MUST clauses; else Required(disjunction, m);
//   * no MUST at all and Optional SHOULD -> the union is what matches (:354-366); Optional + MUST -> RequiredOptionalScorer:
//     score = req + opt when the doc is in opt (reqopt_scorer.rs:78-94); Required + MUST -> Intersection of the two (:405-414);
//   * MUST scorers intersect in ascending cost order, score = their sum (intersection.rs:20-57,325-329); a `+(b OR c)` clause is
//     a union scorer: its score is the sum of its matching clauses (SumCombiner);
//   * MUST_NOT: Exclude (:416-430).
// Exhaustive, dense per-doc evaluation (test sizes).  f32 sums in a fixed order: groups by ascending cost (ties: first clause),
// inside a group / among the SHOULD clauses by descending weight (ties: clause order) -- the canonical order of DESIGN.md §5.
static void bool_for_each(const tqo_index& ix, const tq_query& q, uint32_t segment_ord, const std::function<void(uint32_t, Score)>& emit) {
  if (!q.term_occur) throw std::runtime_error("TQ_OP_BOOL needs term_occur");
  std::vector<const tq_term_seg*> per_term(q.n_terms, nullptr);
  const OSegment* seg = nullptr;
  for (uint32_t i = 0; i < q.n_term_segs; ++i) {
    const tq_term_seg& ts = q.term_segs[i];
    if (ts.segment_ord != segment_ord) continue;
    per_term[ts.term_idx] = &ts;
    if (!seg) seg = &ix.segs.at({ts.segment_ord, ts.field});
  }
  if (!seg) return;
  struct Group { uint32_t id, first_term; uint64_t cost = 0; std::vector<TermScorer> clauses; };
  std::vector<Group> groups;
  std::vector<TermScorer> should, must_not;
  auto scorer_of = [&](uint32_t t) {
    const OSegment& s = ix.segs.at({per_term[t]->segment_ord, per_term[t]->field});
    return make_term_scorer(s, *per_term[t], weight_for(q, t), t, q.term_flags && (q.term_flags[t] & TQ_TERM_IGNORE_FREQ));
  };
  for (uint32_t t = 0; t < q.n_terms; ++t) {
    const bool present = per_term[t] && per_term[t]->doc_freq;
    if (q.term_occur[t] == TQ_OCCUR_MUST) {
      const uint32_t id = q.term_group ? q.term_group[t] : 256u + t;
      Group* g = nullptr;
      for (auto& x : groups) if (x.id == id) g = &x;
      if (!g) { groups.push_back(Group{id, t}); g = &groups.back(); }
      if (present) { g->clauses.push_back(scorer_of(t)); g->cost += per_term[t]->doc_freq; }
    } else if (present) {
      (q.term_occur[t] == TQ_OCCUR_SHOULD ? should : must_not).push_back(scorer_of(t));
    }
  }
  for (auto& g : groups) if (g.clauses.empty()) return;  // an empty MUST
  uint32_t m = q.min_should_match;
  if (m > should.size()) return;
  if (m >= 2 && m == should.size()) {  // as many as there are: they are MUST clauses
    for (auto& sc : should) { Group g{512u + sc.clause, sc.clause}; g.cost = sc.postings.size_hint(); g.clauses.push_back(std::move(sc)); groups.push_back(std::move(g)); }
    should.clear();
    m = 0;
  }
  if (groups.empty() && should.empty()) return;
  const uint32_t need_should = m >= 1 ? m : (groups.empty() ? 1u : 0u);
  const uint32_t max_doc = seg->max_doc;
  auto by_weight = [](std::vector<TermScorer>& v) {
    std::stable_sort(v.begin(), v.end(), [](const TermScorer& a, const TermScorer& b) { return a.similarity_weight.weight > b.similarity_weight.weight; });
  };
  std::stable_sort(groups.begin(), groups.end(), [](const Group& a, const Group& b) { return a.cost < b.cost; });
  std::vector<float> acc(max_doc, 0.0f), gsum(max_doc), ssum(max_doc, 0.0f);
  std::vector<uint8_t> ok(max_doc, 1), gm(max_doc), scount(max_doc, 0);
  bool first = true;
  for (auto& g : groups) {
    by_weight(g.clauses);
    std::fill(gm.begin(), gm.end(), 0);
    for (auto& sc : g.clauses)
      for (uint32_t d = sc.doc(); d != TERMINATED; d = sc.advance()) {
        const Score s = sc.score();
        if (!gm[d]) { gsum[d] = s; gm[d] = 1; } else gsum[d] += s;
      }
    for (uint32_t d = 0; d < max_doc; ++d) {
      if (!ok[d]) continue;
      if (!gm[d]) { ok[d] = 0; continue; }
      acc[d] = first ? gsum[d] : acc[d] + gsum[d];
    }
    first = false;
  }
  by_weight(should);
  for (auto& sc : should)
    for (uint32_t d = sc.doc(); d != TERMINATED; d = sc.advance()) {
      const Score s = sc.score();
      if (!scount[d]) ssum[d] = s; else ssum[d] += s;
      if (scount[d] < 255) ++scount[d];
    }
  for (auto& sc : must_not)
    for (uint32_t d = sc.doc(); d != TERMINATED; d = sc.advance()) ok[d] = 0;
  for (uint32_t d = 0; d < max_doc; ++d) {
    if (!ok[d] || scount[d] < need_should) continue;
    if (groups.empty() && !scount[d]) continue;
    const Score total = groups.empty() ? ssum[d] : (scount[d] ? acc[d] + ssum[d] : acc[d]);
    emit(d, total);
  }
}

// mode 0: exhaustive canonical; mode 1: reference-faithful pruned path.
static void collect_segment(const tqo_index& ix, const tq_query& q, uint32_t segment_ord, int mode, std::vector<Hit>& fruit) {
  // gather this segment's lists per clause
  std::vector<const tq_term_seg*> per_term(q.n_terms, nullptr);
  const OSegment* seg = nullptr;
  for (uint32_t i = 0; i < q.n_term_segs; ++i) {
    const tq_term_seg& ts = q.term_segs[i];
    if (ts.segment_ord != segment_ord) continue;
    per_term[ts.term_idx] = &ts;
    auto it = ix.segs.find({ts.segment_ord, ts.field});
    if (it == ix.segs.end()) throw std::runtime_error("segment/field not registered");
    if (!seg) seg = &it->second;
  }
  if (!seg) return;
  if (q.op == TQ_OP_BOOL) {  // (no pruned variant restated: both modes evaluate exhaustively)
    TopNHeap top_b(q.k);
    const bool has_thr = (q.flags & TQ_QUERY_HAS_THRESHOLD) && q.threshold == q.threshold;
    bool_for_each(ix, q, segment_ord, [&](uint32_t d, Score score) {
      if (!seg->is_alive(d)) return;
      if (has_thr && !(score > q.threshold)) return;
      top_b.push(score, d);
    });
    for (const ScoreHeapEntry& e : top_b.heap) fruit.push_back({e.score, segment_ord, e.doc});
    return;
  }
  if (q.op == TQ_OP_PHRASE) {
    // PhraseWeight::scorer + the same collector (phrase_weight.rs:41-110): every term of the phrase must have postings here
    if (!q.term_pos || !q.term_offset) throw std::runtime_error("phrase query arrays");
    std::vector<std::pair<uint32_t, PhraseTerm>> tp;
    for (uint32_t t = 0; t < q.n_terms; ++t) {
      if (!per_term[t] || per_term[t]->doc_freq == 0) return;
      const tq_term_seg& ts = *per_term[t];
      const tq_term_pos& ps = q.term_pos[per_term[t] - q.term_segs];
      const OSegment& s = ix.segs.at({ts.segment_ord, ts.field});
      if (s.record_option != WithFreqsAndPositions) throw std::runtime_error("field has no positions");
      PhraseTerm pt;
      pt.postings.block_cursor = BlockSegmentPostings::open(ts.doc_freq, s.idx_body.data() + 8 + ts.postings_start,
                                                            (size_t)(ts.postings_end - ts.postings_start), s.record_option, WithFreqsAndPositions);
      pt.postings.cur = 0;
      if (ps.positions_end > s.positions.size() || !PositionReader::open(s.positions.data() + ps.positions_start,
                                                                         (size_t)(ps.positions_end - ps.positions_start), &pt.reader))
        throw std::runtime_error("corrupt positions range");
      tp.emplace_back(q.term_offset[t], std::move(pt));
    }
    PhraseScorer sc;
    sc.slop = q.slop;  // PhraseWeight hands the query's slop to the scorer (phrase_weight.rs:41-110)
    sc.fieldnorm_reader = seg->has_fieldnorm ? FieldNormReader::from_data(seg->fieldnorm.data(), seg->max_doc) : FieldNormReader::constant(seg->max_doc, 1);
    sc.similarity_weight = weight_for(q, 0);  // ONE Bm25Weight::for_terms weight for the whole phrase
    sc.init(std::move(tp));
    TopNHeap top_p(q.k);
    const bool has_thr = (q.flags & TQ_QUERY_HAS_THRESHOLD) && q.threshold == q.threshold;
    for (uint32_t d = sc.doc(); d != TERMINATED; d = sc.advance()) {
      if (!seg->is_alive(d)) continue;
      const Score score = sc.score();
      if (has_thr && !(score > q.threshold)) continue;
      top_p.push(score, d);
    }
    for (const ScoreHeapEntry& e : top_p.heap) fruit.push_back({e.score, segment_ord, e.doc});
    return;
  }
  std::vector<TermScorer> scorers;
  for (uint32_t t = 0; t < q.n_terms; ++t) {
    if (!per_term[t] || per_term[t]->doc_freq == 0) {
      if (q.op == TQ_OP_AND || q.op == TQ_OP_TERM) return;  // Empty scorer => empty intersection
      continue;
    }
    const OSegment& s = ix.segs.at({per_term[t]->segment_ord, per_term[t]->field});
    scorers.push_back(make_term_scorer(s, *per_term[t], weight_for(q, t), t, q.term_flags && (q.term_flags[t] & TQ_TERM_IGNORE_FREQ)));
  }
  if (scorers.empty()) return;
  TopNHeap top_n(q.k);
  const bool has_threshold = (q.flags & TQ_QUERY_HAS_THRESHOLD) && q.threshold == q.threshold;
  auto push = [&](uint32_t doc, Score score) {
    if (!seg->is_alive(doc)) return;
    if (has_threshold && !(score > q.threshold)) return;  // for_each_pruning(threshold, ..): only scores above it (weight.rs:123-132)
    top_n.push(score, doc);
  };
  if (mode == 0) {
    if (scorers.size() == 1) {
      TermScorer& s = scorers[0];
      for (uint32_t d = s.doc(); d != TERMINATED; d = s.advance()) push(d, s.score());
    } else if (q.op == TQ_OP_AND) {
      for_each_intersection(scorers, push);
    } else {
      // canonical union order: descending Bm25Weight.weight, ties in clause order (the reference's order is data
      // dependent: block_wand_union.rs:205-208, buffered_union.rs:69-85) -- SumCombiner then adds in this order
      std::stable_sort(scorers.begin(), scorers.end(), [](const TermScorer& a, const TermScorer& b) { return a.similarity_weight.weight > b.similarity_weight.weight; });
      for_each_union(scorers, push);
    }
  } else {
    Score threshold = has_threshold ? q.threshold : std::numeric_limits<Score>::lowest();
    const Score initial_threshold = threshold;
    PruningCallback cb = [&](uint32_t doc, Score score) -> Score {
      if (!seg->is_alive(doc)) return threshold;  // sort_by_score.rs:44-53
      top_n.push(score, doc);
      threshold = std::max(initial_threshold, top_n.threshold_or_min());
      return threshold;
    };
    if (scorers.size() == 1) block_wand_single_scorer(scorers[0], threshold, cb);
    else if (q.op == TQ_OP_AND) block_wand_intersection(scorers, threshold, cb);
    else block_wand(scorers, threshold, cb);
  }
  for (const ScoreHeapEntry& e : top_n.heap) fruit.push_back({e.score, segment_ord, e.doc});
}

static void search_one(const tqo_index& ix, const tq_query& q, int mode, uint32_t out_stride, float* scores,
                       uint32_t* segs, uint32_t* docs, uint32_t* count) {
  std::vector<uint32_t> seg_ords;
  for (uint32_t i = 0; i < q.n_term_segs; ++i) seg_ords.push_back(q.term_segs[i].segment_ord);
  std::sort(seg_ords.begin(), seg_ords.end());
  seg_ords.erase(std::unique(seg_ords.begin(), seg_ords.end()), seg_ords.end());
  std::vector<Hit> flattened;
  for (uint32_t so : seg_ords) {
    std::vector<Hit> fruit;
    collect_segment(ix, q, so, mode, fruit);
    if (mode == 0) std::sort(fruit.begin(), fruit.end(), hit_before);
    flattened.insert(flattened.end(), fruit.begin(), fruit.end());
  }
  std::vector<Hit> top;
  if (mode == 0) {
    std::sort(flattened.begin(), flattened.end(), hit_before);
    if (flattened.size() > q.k) flattened.resize(q.k);
    top = flattened;
  } else {
    top = merge_top_k(flattened, 0, q.k);
  }
  *count = (uint32_t)top.size();
  for (size_t i = 0; i < top.size() && i < out_stride; ++i) { scores[i] = top[i].score; segs[i] = top[i].segment_ord; docs[i] = top[i].doc; }
}

extern "C" {

tqo_index* tqo_index_create() { return new tqo_index(); }
void tqo_index_destroy(tqo_index* ix) { delete ix; }
const char* tqo_last_error(tqo_index* ix) { return ix->err.c_str(); }

int tqo_segment_register(tqo_index* ix, uint32_t segment_ord, uint32_t field, uint32_t max_doc, int record_option,
                         const uint8_t* idx_body, size_t idx_len, const uint8_t* fieldnorm, size_t fieldnorm_len,
                         const uint8_t* alive, size_t alive_len) {
  OSegment s;
  s.segment_ord = segment_ord; s.field = field; s.max_doc = max_doc; s.record_option = (IndexRecordOption)record_option;
  s.idx_body.assign(idx_body, idx_body + idx_len);
  s.idx_body.resize(idx_len + 64, 0);  // slack for vector over-reads
  s.has_fieldnorm = fieldnorm != nullptr;
  if (fieldnorm) s.fieldnorm.assign(fieldnorm, fieldnorm + fieldnorm_len);
  s.has_alive = alive != nullptr;
  if (alive) s.alive.assign(alive, alive + alive_len);
  ix->segs[{segment_ord, field}] = std::move(s);
  return TQ_OK;
}

// mode: 0 exhaustive canonical top-k, 1 reference-faithful (Block-WAND + TopNHeap + merge_top_k).
// n_threads: independent queries spread over host threads (BASELINE.md threading mode 3).
int tqo_search_batch(tqo_index* ix, const tq_query* queries, size_t nq, int mode, int n_threads, uint32_t out_stride,
                     float* out_scores, uint32_t* out_seg, uint32_t* out_doc, uint32_t* out_count) {
  try {
    std::atomic<size_t> next{0};
    std::atomic<int> failed{0};
    auto work = [&]() {
      for (;;) {
        size_t i = next.fetch_add(1);
        if (i >= nq) break;
        try {
          search_one(*ix, queries[i], mode, out_stride, out_scores + i * out_stride, out_seg + i * out_stride,
                     out_doc + i * out_stride, out_count + i);
        } catch (...) { failed = 1; }
      }
    };
    if (n_threads <= 1) work();
    else {
      std::vector<std::thread> th;
      for (int t = 0; t < n_threads; ++t) th.emplace_back(work);
      for (auto& t : th) t.join();
    }
    if (failed) { ix->err = "search failed"; return TQ_ERR_INVALID_ARGUMENT; }
    return TQ_OK;
  } catch (const std::exception& e) { ix->err = e.what(); return TQ_ERR_INVALID_ARGUMENT; }
}

// Count collector (src/collector/count_collector.rs): Weight::count per segment, summed.  Exhaustive iteration of the
// scorers; alive docs only (a term query without deletes would read doc_freq: the same number).
int tqo_count_batch(tqo_index* ix, const tq_query* queries, size_t nq, uint64_t* out_counts) {
  try {
    for (size_t qi = 0; qi < nq; ++qi) {
      const tq_query& q = queries[qi];
      std::vector<uint32_t> seg_ords;
      for (uint32_t i = 0; i < q.n_term_segs; ++i) seg_ords.push_back(q.term_segs[i].segment_ord);
      std::sort(seg_ords.begin(), seg_ords.end());
      seg_ords.erase(std::unique(seg_ords.begin(), seg_ords.end()), seg_ords.end());
      uint64_t total = 0;
      for (uint32_t so : seg_ords) {
        std::vector<const tq_term_seg*> per_term(q.n_terms, nullptr);
        const OSegment* seg = nullptr;
        for (uint32_t i = 0; i < q.n_term_segs; ++i) {
          const tq_term_seg& ts = q.term_segs[i];
          if (ts.segment_ord != so) continue;
          per_term[ts.term_idx] = &ts;
          if (!seg) seg = &ix->segs.at({ts.segment_ord, ts.field});
        }
        if (q.op == TQ_OP_BOOL) {
          if (seg) bool_for_each(*ix, q, so, [&](uint32_t d, Score) { if (seg->is_alive(d)) ++total; });
          continue;
        }
        std::vector<TermScorer> scorers;
        bool empty = false;
        for (uint32_t t = 0; t < q.n_terms; ++t) {
          if (!per_term[t] || per_term[t]->doc_freq == 0) { if (q.op != TQ_OP_OR) empty = true; continue; }
          const OSegment& s = ix->segs.at({per_term[t]->segment_ord, per_term[t]->field});
          Bm25Weight w{};
          w.weight = 1.0f; w.average_fieldnorm = 1.0f;
          for (int id = 0; id < 256; ++id) w.cache[id] = 1.0f;
          scorers.push_back(make_term_scorer(s, *per_term[t], w, t));
        }
        if (empty || scorers.empty()) continue;
        auto hit = [&](uint32_t doc, Score) { if (seg->is_alive(doc)) ++total; };
        if (scorers.size() == 1) {
          TermScorer& s = scorers[0];
          for (uint32_t d = s.doc(); d != TERMINATED; d = s.advance()) hit(d, 0.0f);
        } else if (q.op == TQ_OP_AND) {
          for_each_intersection(scorers, hit);
        } else {
          for_each_union(scorers, hit);
        }
      }
      out_counts[qi] = total;
    }
  } catch (const std::exception& e) { ix->err = e.what(); return TQ_ERR_INVALID_ARGUMENT; }
  return TQ_OK;
}

int tqo_decode_postings(tqo_index* ix, const tq_term_seg* ts, uint32_t* out_docs, uint32_t* out_tfs) {
  try {
    const OSegment& seg = ix->segs.at({ts->segment_ord, ts->field});
    Bm25Weight w; w.weight = 1; for (auto& c : w.cache) c = 1;
    TermScorer t = make_term_scorer(seg, *ts, w, 0);
    uint32_t i = 0;
    for (uint32_t d = t.doc(); d != TERMINATED; d = t.advance()) {
      if (i >= ts->doc_freq) { ix->err = "more docs than doc_freq"; return TQ_ERR_CORRUPT; }
      out_docs[i] = d; if (out_tfs) out_tfs[i] = t.term_freq(); ++i;
    }
    if (i != ts->doc_freq) { ix->err = "fewer docs than doc_freq"; return TQ_ERR_CORRUPT; }
    return TQ_OK;
  } catch (const std::exception& e) { ix->err = e.what(); return TQ_ERR_INVALID_ARGUMENT; }
}

int tqo_block_table(tqo_index* ix, const tq_term_seg* ts, float weight, float avg_fieldnorm, uint32_t* out_last_doc, float* out_block_max) {
  try {
    const OSegment& seg = ix->segs.at({ts->segment_ord, ts->field});
    Bm25Weight w; w.weight = weight; w.average_fieldnorm = avg_fieldnorm;
    for (int id = 0; id < 256; ++id) w.cache[id] = cached_tf_component(id_to_fieldnorm((uint8_t)id), avg_fieldnorm);
    TermScorer t = make_term_scorer(seg, *ts, w, 0);
    const uint32_t n = ts->doc_freq / COMPRESSION_BLOCK_SIZE;
    SkipReader& sr = t.postings.block_cursor.skip_reader;
    for (uint32_t b = 0; b < n; ++b) {
      out_last_doc[b] = sr.last_doc_in_block();
      Score s = 0; sr.block_max_score(w, &s); out_block_max[b] = s;
      sr.advance();
    }
    return TQ_OK;
  } catch (const std::exception& e) { ix->err = e.what(); return TQ_ERR_INVALID_ARGUMENT; }
}

// ---- TermScorer handle, driven by tests exactly like the reference's unit tests drive it -------
struct tqo_term_scorer { TermScorer s; };
tqo_term_scorer* tqo_term_scorer_open(tqo_index* ix, const tq_term_seg* ts, float weight, float avg_fieldnorm) {
  try {
    const OSegment& seg = ix->segs.at({ts->segment_ord, ts->field});
    Bm25Weight w; w.weight = weight; w.average_fieldnorm = avg_fieldnorm;
    for (int id = 0; id < 256; ++id) w.cache[id] = cached_tf_component(id_to_fieldnorm((uint8_t)id), avg_fieldnorm);
    auto* h = new tqo_term_scorer{make_term_scorer(seg, *ts, w, 0)};
    return h;
  } catch (const std::exception& e) { ix->err = e.what(); return nullptr; }
}
void tqo_term_scorer_close(tqo_term_scorer* h) { delete h; }
uint32_t tqo_term_scorer_doc(tqo_term_scorer* h) { return h->s.doc(); }
uint32_t tqo_term_scorer_advance(tqo_term_scorer* h) { return h->s.advance(); }
uint32_t tqo_term_scorer_seek(tqo_term_scorer* h, uint32_t target) { return h->s.seek(target); }
void tqo_term_scorer_seek_block(tqo_term_scorer* h, uint32_t target) { h->s.seek_block(target); }
uint32_t tqo_term_scorer_term_freq(tqo_term_scorer* h) { return h->s.term_freq(); }
uint32_t tqo_term_scorer_last_doc_in_block(tqo_term_scorer* h) { return h->s.last_doc_in_block(); }
float tqo_term_scorer_score(tqo_term_scorer* h) { return h->s.score(); }
float tqo_term_scorer_block_max_score(tqo_term_scorer* h) { return h->s.block_max_score(); }
float tqo_term_scorer_max_score(tqo_term_scorer* h) { return h->s.max_score(); }

// ---- codec hooks ----------------------------------------------------------------------------
uint8_t tqo_bp4x_num_bits(const uint32_t* v) { return bp4x_num_bits(v); }
void tqo_bp4x_pack(const uint32_t* v, uint8_t b, uint8_t* out) { bp4x_pack(v, b, out); }
void tqo_bp4x_unpack(const uint8_t* in, uint8_t b, uint32_t* out) { bp4x_unpack(in, b, out); }
void tqo_bp4x_unpack_scalar(const uint8_t* in, uint8_t b, uint32_t* out) { bp4x_unpack_scalar(in, b, out); }
// BlockEncoder::compress_block_sorted / BlockDecoder::uncompress_block_sorted
uint8_t tqo_compress_block_sorted(const uint32_t* block, uint32_t offset, uint8_t* out, size_t* written) {
  BlockEncoder e; uint8_t nb = e.compress_block_sorted(block, offset, written); std::memcpy(out, e.output, *written); return nb;
}
uint8_t tqo_compress_block_unsorted(const uint32_t* block, int minus_one, uint8_t* out, size_t* written) {
  BlockEncoder e; uint8_t nb = e.compress_block_unsorted(block, minus_one != 0, written); std::memcpy(out, e.output, *written); return nb;
}
size_t tqo_uncompress_block_sorted(const uint8_t* data, uint32_t offset, uint8_t num_bits, int strict, uint32_t* out) {
  std::vector<uint8_t> padded(data, data + compressed_block_size(num_bits)); padded.resize(padded.size() + 32);
  BlockDecoder d; size_t n = d.uncompress_block_sorted(padded.data(), offset, num_bits, strict != 0); std::memcpy(out, d.output, 512); return n;
}
size_t tqo_uncompress_block_unsorted(const uint8_t* data, uint8_t num_bits, int minus_one, uint32_t* out) {
  std::vector<uint8_t> padded(data, data + compressed_block_size(num_bits)); padded.resize(padded.size() + 32);
  BlockDecoder d; size_t n = d.uncompress_block_unsorted(padded.data(), num_bits, minus_one != 0); std::memcpy(out, d.output, 512); return n;
}
size_t tqo_vint_compress_sorted(const uint32_t* in, size_t n, uint32_t offset, uint8_t* out) { return vint_compress_sorted(in, n, out, offset); }
size_t tqo_vint_compress_unsorted(const uint32_t* in, size_t n, uint8_t* out) { return vint_compress_unsorted(in, n, out); }
size_t tqo_vint_uncompress_sorted(const uint8_t* data, size_t n, uint32_t offset, uint32_t padding, uint32_t* out128) {
  BlockDecoder d; size_t r = d.uncompress_vint_sorted(data, offset, n, padding); std::memcpy(out128, d.output, 512); return r;
}
size_t tqo_search_block(const uint32_t* arr128, uint32_t target) { return search_block(arr128, target); }
uint8_t tqo_encode_bitwidth(uint8_t bw, int delta1) { return encode_bitwidth(bw, delta1 != 0); }
uint8_t tqo_encode_block_wand_max_tf(uint32_t tf) { return encode_block_wand_max_tf(tf); }
uint32_t tqo_decode_block_wand_max_tf(uint8_t c) { return decode_block_wand_max_tf(c); }

// ---- BM25 / fieldnorm hooks -----------------------------------------------------------------
float tqo_bm25_idf(uint64_t doc_freq, uint64_t doc_count) { return idf(doc_freq, doc_count); }
float tqo_bm25_weight(uint64_t doc_freq, uint64_t doc_count, float boost) {
  return Bm25Weight::for_one_term(doc_freq, doc_count, 1.0f).boost_by(boost).weight;
}
void tqo_bm25_tf_cache(float avg, float* out) { for (int id = 0; id < 256; ++id) out[id] = cached_tf_component(id_to_fieldnorm((uint8_t)id), avg); }
uint32_t tqo_id_to_fieldnorm(uint8_t id) { return id_to_fieldnorm(id); }
uint8_t tqo_fieldnorm_to_id(uint32_t f) { return fieldnorm_to_id(f); }

// ---- serializer hooks (PostingsSerializer) ---------------------------------------------------
struct tqo_field_writer {
  std::vector<uint8_t> body; IndexRecordOption mode; std::vector<uint8_t> fieldnorm_ids; bool has_fn; float avg;
};
// ---- phrase queries (N3 groundwork): PhraseScorer over registered segments ------------------------------------------
int tqo_segment_register_positions(tqo_index* ix, uint32_t segment_ord, uint32_t field, const uint8_t* pos, size_t len) {
  auto it = ix->segs.find({segment_ord, field});
  if (it == ix->segs.end()) { ix->err = "segment/field not registered"; return TQ_ERR_NOT_FOUND; }
  it->second.positions.assign(pos, pos + len);
  return TQ_OK;
}

typedef struct {
  uint32_t offset;  // position of the term inside the phrase
  uint32_t segment_ord, field, doc_freq;
  uint64_t postings_start, postings_end, positions_start, positions_end;
} tqo_phrase_term;

// All matches of one phrase in (segment, doc) order: PhraseWeight::scorer + a collector that keeps everything
// (phrase_weight.rs:41-110; weight = Bm25Weight::for_terms over the phrase's terms, supplied by the caller).
int tqo_phrase_search(tqo_index* ix, const tqo_phrase_term* terms, size_t n, uint32_t n_phrase_terms, float weight, float avg_fieldnorm,
                      uint32_t slop, size_t cap, uint32_t* out_seg, uint32_t* out_doc, float* out_score, uint32_t* out_count, size_t* out_n) {
  try {
    std::vector<uint32_t> seg_ords;
    for (size_t i = 0; i < n; ++i) seg_ords.push_back(terms[i].segment_ord);
    std::sort(seg_ords.begin(), seg_ords.end());
    seg_ords.erase(std::unique(seg_ords.begin(), seg_ords.end()), seg_ords.end());
    size_t w = 0;
    for (uint32_t so : seg_ords) {
      std::vector<std::pair<uint32_t, PhraseTerm>> tp;
      const OSegment* seg = nullptr;
      for (size_t i = 0; i < n; ++i) {
        const tqo_phrase_term& t = terms[i];
        if (t.segment_ord != so || t.doc_freq == 0) continue;
        const OSegment& s = ix->segs.at({t.segment_ord, t.field});
        seg = &s;
        if (s.record_option != WithFreqsAndPositions) throw std::runtime_error("field has no positions");
        PhraseTerm pt;
        pt.postings.block_cursor = BlockSegmentPostings::open(t.doc_freq, s.idx_body.data() + 8 + t.postings_start,
                                                              (size_t)(t.postings_end - t.postings_start), s.record_option, WithFreqsAndPositions);
        pt.postings.cur = 0;
        if (t.positions_end > s.positions.size() || !PositionReader::open(s.positions.data() + t.positions_start,
                                                                          (size_t)(t.positions_end - t.positions_start), &pt.reader))
          throw std::runtime_error("corrupt positions range");
        tp.emplace_back(t.offset, std::move(pt));
      }
      if (!seg || tp.size() < n_phrase_terms) continue;  // a term without postings in this segment: no match
      PhraseScorer sc;
      sc.slop = slop;
      sc.fieldnorm_reader = seg->has_fieldnorm ? FieldNormReader::from_data(seg->fieldnorm.data(), seg->max_doc) : FieldNormReader::constant(seg->max_doc, 1);
      sc.similarity_weight.weight = weight;
      sc.similarity_weight.average_fieldnorm = avg_fieldnorm;
      for (int id = 0; id < 256; ++id) sc.similarity_weight.cache[id] = cached_tf_component(id_to_fieldnorm((uint8_t)id), avg_fieldnorm);
      sc.init(std::move(tp));
      for (uint32_t d = sc.doc(); d != TERMINATED; d = sc.advance()) {
        if (!seg->is_alive(d)) continue;
        if (w < cap) { out_seg[w] = so; out_doc[w] = d; out_score[w] = sc.score(); out_count[w] = sc.phrase_count(); }
        ++w;
      }
    }
    *out_n = w;
  } catch (const std::exception& e) { ix->err = e.what(); return TQ_ERR_INVALID_ARGUMENT; }
  return TQ_OK;
}

// ---- positions codec (N3 groundwork) ---------------------------------------------------------------------------
// One term: serialises `n` position deltas handed over in chunks of `chunk` (write_positions_delta may be called
// several times per term); returns the byte length, bytes in *out (malloc'd by the caller via two calls: out == NULL
// just sizes).
size_t tqo_positions_serialize(const uint32_t* deltas, size_t n, size_t chunk, uint8_t* out, size_t out_cap) {
  PositionSerializer ser;
  if (chunk == 0) chunk = n ? n : 1;
  for (size_t i = 0; i < n; i += chunk) ser.write_positions_delta(deltas + i, std::min(chunk, n - i));
  ser.close_term();
  if (out && out_cap >= ser.out.size()) std::memcpy(out, ser.out.data(), ser.out.size());
  return ser.out.size();
}
struct tqo_position_reader { std::vector<uint8_t> data; PositionReader r; };
tqo_position_reader* tqo_position_reader_open(const uint8_t* data, size_t len) {
  auto* h = new tqo_position_reader();
  h->data.assign(data, data + len);
  if (!PositionReader::open(h->data.data(), h->data.size(), &h->r)) { delete h; return nullptr; }
  return h;
}
void tqo_position_reader_read(tqo_position_reader* h, uint64_t offset, uint32_t* out, size_t n) { h->r.read(offset, out, n); }
void tqo_position_reader_close(tqo_position_reader* h) { delete h; }

tqo_field_writer* tqo_field_writer_create(int record_option, uint64_t total_num_tokens, const uint8_t* fieldnorm_ids, uint32_t max_doc) {
  auto* w = new tqo_field_writer();
  w->mode = (IndexRecordOption)record_option; w->has_fn = fieldnorm_ids != nullptr;
  if (fieldnorm_ids) w->fieldnorm_ids.assign(fieldnorm_ids, fieldnorm_ids + max_doc);
  // FieldSerializer::create (serializer.rs:120-133): header + segment-local average fieldnorm
  uint8_t hdr[8]; std::memcpy(hdr, &total_num_tokens, 8); w->body.assign(hdr, hdr + 8);
  w->avg = w->has_fn ? (float)total_num_tokens / (float)max_doc : 0.0f;
  return w;
}
int tqo_field_writer_add_term(tqo_field_writer* w, const uint32_t* docs, const uint32_t* tfs, uint32_t doc_freq, uint64_t* start, uint64_t* end) {
  FieldNormReader fnr = FieldNormReader::from_data(w->fieldnorm_ids.data(), (uint32_t)w->fieldnorm_ids.size());
  PostingsSerializer ser(w->avg, w->mode, w->has_fn ? &fnr : nullptr);
  try { ser.new_term(doc_freq, tfs != nullptr); } catch (...) { return TQ_ERR_INVALID_ARGUMENT; }
  for (uint32_t i = 0; i < doc_freq; ++i) ser.write_doc(docs[i], tfs ? tfs[i] : 1u);
  *start = w->body.size() - 8;
  ser.close_term(doc_freq, w->body);
  *end = w->body.size() - 8;
  return TQ_OK;
}
int tqo_field_writer_body(tqo_field_writer* w, const uint8_t** body, size_t* len) { *body = w->body.data(); *len = w->body.size(); return TQ_OK; }
void tqo_field_writer_destroy(tqo_field_writer* w) { delete w; }

// ---- collector hooks --------------------------------------------------------------------------
// Feeds (score, doc) pairs in order into TopNHeap; writes the threshold after every push
// (NaN-free; -inf encodes None) and the final content sorted (score desc, doc asc).
size_t tqo_top_n_heap(const float* scores, const uint32_t* docs, size_t n, size_t k, float* thresholds, float* out_scores, uint32_t* out_docs) {
  TopNHeap h(k);
  for (size_t i = 0; i < n; ++i) { h.push(scores[i], docs[i]); if (thresholds) thresholds[i] = h.has_threshold ? h.threshold : -INFINITY; }
  std::vector<Hit> v; for (auto& e : h.heap) v.push_back({e.score, 0, e.doc});
  std::sort(v.begin(), v.end(), hit_before);
  for (size_t i = 0; i < v.size(); ++i) { out_scores[i] = v[i].score; out_docs[i] = v[i].doc; }
  return v.size();
}
size_t tqo_merge_top_k(const float* scores, const uint32_t* segs, const uint32_t* docs, size_t n, size_t start, size_t end,
                       float* out_scores, uint32_t* out_segs, uint32_t* out_docs) {
  std::vector<Hit> f; for (size_t i = 0; i < n; ++i) f.push_back({scores[i], segs[i], docs[i]});
  std::vector<Hit> r = merge_top_k(f, start, end);
  for (size_t i = 0; i < r.size(); ++i) { out_scores[i] = r[i].score; out_segs[i] = r[i].segment_ord; out_docs[i] = r[i].doc; }
  return r.size();
}

}  // extern "C"
