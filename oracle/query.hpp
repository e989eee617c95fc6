// ORACLE — TEST INFRASTRUCTURE ONLY (see codec.hpp header).
// CPU restatement of the reference's scorers, Block-WAND algorithms and top-k collectors.
// Follows (relative to /root/reference):
//   src/query/term_query/term_scorer.rs:9-150                TermScorer
//   src/query/boolean_query/block_wand_union.rs:16-265       block_wand, block_wand_single_scorer
//   src/query/boolean_query/block_wand_intersection.rs:19-179 block_wand_intersection
//   src/query/intersection.rs:122-179,325-329                exhaustive leap-frog AND (semantics)
//   src/query/union/buffered_union.rs:63-151                 exhaustive OR (semantics)
//   src/collector/sort_key/sort_by_score.rs:85-161           ScoreHeapEntry, TopNHeap
//   src/collector/top_score_collector.rs:512-662             TopNComputer, compare_for_top_k
//   src/collector/sort_key_top_collector.rs:76-95            merge_top_k
#pragma once
#include <functional>
#include <limits>

#include "postings.hpp"

namespace tqo {

// ---- TermScorer ---------------------------------------------------------------------------
struct TermScorer {
  SegmentPostings postings;
  FieldNormReader fieldnorm_reader;
  Bm25Weight similarity_weight;
  uint32_t clause = 0;  // position of the term in the query (not in the reference; used only by
                        // the canonical fixed-order exhaustive scorers below)

  void seek_block(uint32_t target) { postings.block_cursor.seek_block(target); }
  Score block_max_score() { return postings.block_cursor.block_max_score(fieldnorm_reader, similarity_weight); }
  uint32_t term_freq() const { return postings.term_freq(); }
  uint8_t fieldnorm_id() const { return fieldnorm_reader.fieldnorm_id(doc()); }
  Score max_score() const { return similarity_weight.max_score(); }
  uint32_t last_doc_in_block() const { return postings.block_cursor.skip_reader.last_doc_in_block(); }
  BlockSegmentPostings& block_cursor() { return postings.block_cursor; }
  uint32_t advance() { return postings.advance(); }
  uint32_t seek(uint32_t target) { return postings.seek(target); }
  uint32_t doc() const { return postings.doc(); }
  uint32_t size_hint() const { return postings.size_hint(); }
  Score score() const { return similarity_weight.score(fieldnorm_id(), term_freq()); }
};

typedef std::function<Score(uint32_t, Score)> PruningCallback;

// ---- block_wand_single_scorer (block_wand_union.rs:226-265) -------------------------------
inline void block_wand_single_scorer(TermScorer scorer, Score threshold, const PruningCallback& callback) {
  uint32_t doc = scorer.doc();
  for (;;) {
    while (scorer.block_max_score() <= threshold) {
      const uint32_t last_doc_in_block = scorer.last_doc_in_block();
      if (last_doc_in_block == TERMINATED) return;
      doc = last_doc_in_block + 1;
      scorer.seek_block(doc);
    }
    doc = scorer.seek(doc);
    if (doc == TERMINATED) break;
    for (;;) {
      const Score score = scorer.score();
      if (score > threshold) threshold = callback(doc, score);
      if (doc == scorer.last_doc_in_block()) break;
      doc = scorer.advance();
      if (doc == TERMINATED) return;
    }
    doc += 1;
    scorer.seek_block(doc);
  }
}

// ---- block_wand (block_wand_union.rs:16-216) ------------------------------------------------
struct TermScorerWithMaxScore { TermScorer* scorer; Score max_score; };
typedef std::vector<TermScorerWithMaxScore> WandScorers;

inline bool find_pivot_doc(const WandScorers& ts, Score threshold, size_t* before_pivot_len, size_t* pivot_len, uint32_t* pivot_doc) {
  Score max_score = 0.0f;
  size_t before = 0;
  uint32_t pivot = TERMINATED;
  while (before < ts.size()) {
    max_score += ts[before].max_score;
    if (max_score > threshold) { pivot = ts[before].scorer->doc(); break; }
    before += 1;
  }
  if (pivot == TERMINATED) return false;
  size_t plen = before + 1;
  while (plen < ts.size() && ts[plen].scorer->doc() == pivot) plen += 1;
  *before_pivot_len = before; *pivot_len = plen; *pivot_doc = pivot;
  return true;
}
inline void restore_ordering(WandScorers& ts, size_t ord) {
  const uint32_t doc = ts[ord].scorer->doc();
  for (size_t i = ord + 1; i < ts.size(); ++i) {
    if (ts[i].scorer->doc() >= doc) break;
    std::swap(ts[i], ts[i - 1]);
  }
}
inline void block_max_was_too_low_advance_one_scorer(WandScorers& scorers, size_t pivot_len) {
  size_t scorer_to_seek = pivot_len - 1;
  Score global_max_score = scorers[scorer_to_seek].max_score;
  uint32_t doc_to_seek_after = scorers[scorer_to_seek].scorer->last_doc_in_block();
  for (size_t ord = pivot_len - 1; ord-- > 0;) {
    TermScorer* s = scorers[ord].scorer;
    if (s->last_doc_in_block() <= doc_to_seek_after) doc_to_seek_after = s->last_doc_in_block();
    if (scorers[ord].max_score > global_max_score) { global_max_score = scorers[ord].max_score; scorer_to_seek = ord; }
  }
  if (doc_to_seek_after != TERMINATED) doc_to_seek_after += 1;
  for (size_t i = pivot_len; i < scorers.size(); ++i)
    if (scorers[i].scorer->doc() <= doc_to_seek_after) doc_to_seek_after = scorers[i].scorer->doc();
  scorers[scorer_to_seek].scorer->seek(doc_to_seek_after);
  restore_ordering(scorers, scorer_to_seek);
}
inline void swap_remove(WandScorers& v, size_t i) { v[i] = v.back(); v.pop_back(); }
inline bool align_scorers(WandScorers& ts, uint32_t pivot_doc, size_t before_pivot_len) {
  for (size_t i = before_pivot_len; i-- > 0;) {
    const uint32_t new_doc = ts[i].scorer->seek(pivot_doc);
    if (new_doc != pivot_doc) {
      if (new_doc == TERMINATED) swap_remove(ts, i);
      // After swap_remove the element now at `i` may be the former last one; the reference
      // calls restore_ordering(term_scorers, i) unconditionally (block_wand_union.rs:119).
      if (i < ts.size()) restore_ordering(ts, i);
      return false;
    }
  }
  return true;
}
inline void advance_all_scorers_on_pivot(WandScorers& ts, size_t pivot_len) {
  for (size_t i = 0; i < pivot_len; ++i) ts[i].scorer->advance();
  size_t i = 0;
  while (i != ts.size()) {
    if (ts[i].scorer->doc() == TERMINATED) swap_remove(ts, i); else i += 1;
  }
  std::stable_sort(ts.begin(), ts.end(), [](const TermScorerWithMaxScore& a, const TermScorerWithMaxScore& b) { return a.scorer->doc() < b.scorer->doc(); });
}
inline void block_wand(std::vector<TermScorer>& scorers_in, Score threshold, const PruningCallback& callback) {
  std::vector<TermScorer*> live;
  for (auto& s : scorers_in) if (s.doc() < TERMINATED) live.push_back(&s);
  if (live.size() == 1) { block_wand_single_scorer(*live[0], threshold, callback); return; }
  WandScorers scorers;
  for (auto* s : live) scorers.push_back({s, s->max_score()});
  std::stable_sort(scorers.begin(), scorers.end(), [](const TermScorerWithMaxScore& a, const TermScorerWithMaxScore& b) { return a.scorer->doc() < b.scorer->doc(); });
  size_t before_pivot_len, pivot_len; uint32_t pivot_doc;
  while (find_pivot_doc(scorers, threshold, &before_pivot_len, &pivot_len, &pivot_doc)) {
    Score block_max_score_upperbound = 0.0f;  // Iterator::sum for f32 starts from 0.0 (-0.0 in newer std; irrelevant for positives)
    for (size_t i = 0; i < pivot_len; ++i) {
      scorers[i].scorer->seek_block(pivot_doc);
      block_max_score_upperbound += scorers[i].scorer->block_max_score();
    }
    if (block_max_score_upperbound <= threshold) { block_max_was_too_low_advance_one_scorer(scorers, pivot_len); continue; }
    if (!align_scorers(scorers, pivot_doc, before_pivot_len)) continue;
    Score score = 0.0f;
    for (size_t i = 0; i < pivot_len; ++i) score += scorers[i].scorer->score();
    if (score > threshold) threshold = callback(pivot_doc, score);
    advance_all_scorers_on_pivot(scorers, pivot_len);
  }
}

// ---- block_wand_intersection (block_wand_intersection.rs:19-179) -----------------------------
inline void block_wand_intersection(std::vector<TermScorer>& scorers, Score threshold, const PruningCallback& callback) {
  assert(scorers.size() >= 2);
  std::stable_sort(scorers.begin(), scorers.end(), [](const TermScorer& a, const TermScorer& b) { return a.size_hint() < b.size_hint(); });
  TermScorer& leader = scorers[0];
  const size_t num_secondaries = scorers.size() - 1;
  TermScorer* secondaries = &scorers[1];
  const Score leader_max_score = leader.max_score();
  Score secondaries_global_max_sum = 0.0f;
  for (size_t i = 0; i < num_secondaries; ++i) secondaries_global_max_sum += secondaries[i].max_score();
  if (leader_max_score + secondaries_global_max_sum <= threshold) return;
  const FieldNormReader fieldnorm_reader = leader.fieldnorm_reader;
  const Bm25Weight bm25_weight = leader.similarity_weight;
  uint32_t doc = leader.doc();
  std::vector<float> secondary_block_max_scores(num_secondaries, 0.0f), secondary_suffix_block_max(num_secondaries, 0.0f);
  while (doc < TERMINATED) {
    leader.seek_block(doc);
    const Score leader_block_max = leader.block_max_score();
    uint32_t window_end = leader.last_doc_in_block();
    Score secondary_block_max_sum = 0.0f;
    for (size_t idx = 0; idx < num_secondaries; ++idx) {
      TermScorer& secondary = secondaries[idx];
      secondary.block_cursor().seek_block(doc);
      if (!secondary.block_cursor().has_remaining_docs()) return;
      window_end = std::min(window_end, secondary.last_doc_in_block());
      const Score bms = secondary.block_max_score();
      secondary_block_max_scores[idx] = bms;
      secondary_block_max_sum += bms;
    }
    if (leader_block_max + secondary_block_max_sum <= threshold) { doc = window_end + 1; continue; }
    BlockSegmentPostings& block_cursor = leader.block_cursor();
    const size_t start_idx = block_cursor.seek(doc);
    const size_t end_idx = std::min(search_block(block_cursor.doc_decoder.output, window_end + 1), block_cursor.block_len());
    const Score score_threshold = threshold - secondary_block_max_sum;
    uint32_t candidate_doc_ids[COMPRESSION_BLOCK_SIZE + 1]; float candidate_scores[COMPRESSION_BLOCK_SIZE + 1];
    size_t num_candidates = 0;
    for (size_t i = start_idx; i < end_idx; ++i) {
      const uint32_t candidate_doc = block_cursor.doc_decoder.output[i];
      const uint32_t term_freq = block_cursor.freq_decoder.output[i];
      const Score leader_score = bm25_weight.score(fieldnorm_reader.fieldnorm_id(candidate_doc), term_freq);
      candidate_doc_ids[num_candidates] = candidate_doc;
      candidate_scores[num_candidates] = leader_score;
      num_candidates += (leader_score > score_threshold) ? 1 : 0;
    }
    if (num_candidates == 0) { doc = window_end + 1; continue; }
    float running = 0.0f;
    for (size_t idx = num_secondaries; idx-- > 0;) { secondary_suffix_block_max[idx] = running; running += secondary_block_max_scores[idx]; }
    for (size_t ci = 0; ci < num_candidates; ++ci) {
      const uint32_t candidate_doc = candidate_doc_ids[ci];
      Score total_score = candidate_scores[ci];
      bool matched = true;
      for (size_t si = 0; si < num_secondaries; ++si) {
        TermScorer& secondary = secondaries[si];
        if (secondary.doc() > candidate_doc) { matched = false; break; }
        if (secondary.seek(candidate_doc) != candidate_doc) { matched = false; break; }
        total_score += secondary.score();
        if (total_score + secondary_suffix_block_max[si] <= threshold) { matched = false; break; }
      }
      if (!matched) continue;
      if (total_score > threshold) {
        threshold = callback(candidate_doc, total_score);
        if (leader_max_score + secondaries_global_max_sum <= threshold) return;
      }
    }
    doc = window_end + 1;
  }
}

// ---- exhaustive scorers (canonical truth; SURVEY.md §8c caveats (i),(ii)) -------------------
// AND: score = leader + secondaries in ascending-doc_freq (stable) order, i.e. the summation order
// of both Intersection::score (intersection.rs:325-329) and block_wand_intersection (:146-158).
inline void for_each_intersection(std::vector<TermScorer>& scorers, const std::function<void(uint32_t, Score)>& cb) {
  std::stable_sort(scorers.begin(), scorers.end(), [](const TermScorer& a, const TermScorer& b) { return a.size_hint() < b.size_hint(); });
  uint32_t candidate = scorers[0].doc();
  while (candidate < TERMINATED) {
    candidate = scorers[0].seek(candidate);
    if (candidate == TERMINATED) break;
    bool all = true;
    for (size_t i = 1; i < scorers.size(); ++i) {
      const uint32_t d = scorers[i].seek(candidate);
      if (d != candidate) { candidate = d; all = false; break; }
    }
    if (!all) continue;
    Score total = scorers[0].score();
    for (size_t i = 1; i < scorers.size(); ++i) total += scorers[i].score();
    cb(candidate, total);
    candidate += 1;
  }
}
// OR: SumCombiner (score_combiner.rs:39-57) starts at 0.0 and adds every matching scorer's score;
// here always in CLAUSE order (the reference's order is data dependent: buffered_union.rs:69-85,
// block_wand_union.rs:205-208).
inline void for_each_union(std::vector<TermScorer>& scorers, const std::function<void(uint32_t, Score)>& cb) {
  for (;;) {
    uint32_t min_doc = TERMINATED;
    for (auto& s : scorers) min_doc = std::min(min_doc, s.doc());
    if (min_doc == TERMINATED) break;
    Score total = 0.0f;
    for (auto& s : scorers) if (s.doc() == min_doc) { total += s.score(); s.advance(); }
    cb(min_doc, total);
  }
}

// ---- TopNHeap (sort_by_score.rs:85-161) -------------------------------------------------------
struct ScoreHeapEntry { Score score; uint32_t doc; };
// Ord: score asc (partial_cmp, Equal on NaN) then doc DESC  => "greater" = better.
inline bool entry_less(const ScoreHeapEntry& a, const ScoreHeapEntry& b) {
  if (a.score < b.score) return true;
  if (a.score > b.score) return false;
  return a.doc > b.doc;
}
struct TopNHeap {
  std::vector<ScoreHeapEntry> heap;  // min-heap on entry order (BinaryHeap<Reverse<..>>)
  size_t top_n;
  bool has_threshold = false; Score threshold = 0;
  explicit TopNHeap(size_t n) : top_n(n) { heap.reserve(n); }
  static bool cmp(const ScoreHeapEntry& a, const ScoreHeapEntry& b) { return entry_less(b, a); }  // std heap = max-heap of "cmp less"
  void push(Score score, uint32_t doc) {
    if (heap.size() < top_n) {
      heap.push_back({score, doc}); std::push_heap(heap.begin(), heap.end(), cmp);
      if (heap.size() == top_n) { has_threshold = true; threshold = heap.front().score; }
    } else if (has_threshold) {
      if (score > threshold) {
        std::pop_heap(heap.begin(), heap.end(), cmp); heap.back() = {score, doc}; std::push_heap(heap.begin(), heap.end(), cmp);
        threshold = heap.front().score;
      }
    }
  }
  Score threshold_or_min() const { return has_threshold ? threshold : std::numeric_limits<Score>::lowest(); }  // Score::MIN
};

// ---- TopNComputer / merge_top_k (top_score_collector.rs:512-662, sort_key_top_collector.rs:76-95)
struct Hit { Score score; uint32_t segment_ord; uint32_t doc; };
// compare_for_top_k with NaturalComparator on Score: score desc, then DocAddress asc.
inline bool hit_before(const Hit& a, const Hit& b) {
  if (a.score > b.score) return true;
  if (a.score < b.score) return false;
  if (a.segment_ord != b.segment_ord) return a.segment_ord < b.segment_ord;
  return a.doc < b.doc;
}
struct TopNComputer {
  std::vector<Hit> buffer; size_t top_n, cap; bool has_threshold = false; Score threshold = 0;
  explicit TopNComputer(size_t n) : top_n(n), cap(std::max<size_t>(n, 1) * 2) { buffer.reserve(cap); }
  void push(const Hit& h) {
    if (has_threshold && !(h.score > threshold)) return;
    if (buffer.size() == cap) { threshold = truncate_top_n(); has_threshold = true; }
    buffer.push_back(h);
  }
  Score truncate_top_n() {
    std::nth_element(buffer.begin(), buffer.begin() + top_n, buffer.end(), hit_before);
    const Score median = buffer[top_n].score;
    buffer.resize(top_n);
    return median;
  }
  std::vector<Hit> into_sorted_vec() {
    if (buffer.size() > top_n) truncate_top_n();
    std::sort(buffer.begin(), buffer.end(), hit_before);
    return buffer;
  }
};
inline std::vector<Hit> merge_top_k(const std::vector<Hit>& flattened, size_t start, size_t end) {
  if (end <= start) return {};
  TopNComputer c(end);
  for (const Hit& h : flattened) c.push(h);
  std::vector<Hit> sorted = c.into_sorted_vec();
  if (start >= sorted.size()) return {};
  return std::vector<Hit>(sorted.begin() + start, sorted.end());
}

}  // namespace tqo
