// Deterministic synthetic segments in tantivy's on-disk posting format (host, product code).
//
// Workload model of SURVEY.md §8(d), mirroring the shapes of the reference's own bench
// harnesses (benches/and_or_queries.rs:129-155, benches/intersection_bench.rs:20-36,107-113):
//   * doc length  ~ round(lognormal(mu = ln 80, sigma = 0.6)) clipped to [1, 4096] -> fieldnorm id
//   * a term of density p: doc ids by geometric gap sampling (mean gap 1/p)
//   * tf = 1 + geometric(p = 0.7) capped at 10, clipped to the doc's length
//   * Zipf vocabulary: rank r has density min(0.5, c / r)
// Posting lists are generated directly (no tokenizer / indexer) and serialized with
// FieldPostingsWriter, i.e. byte-for-byte what PostingsSerializer would write
// (src/postings/serializer.rs:353-481).  One RNG stream per (segment, term) and per fieldnorm
// chunk makes the output independent of the thread count.
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdint>
#include <memory>
#include <thread>
#include <vector>

#include "segment_writer.hpp"

namespace {

struct Rng {  // xoshiro256** seeded by splitmix64
  uint64_t s[4];
  static uint64_t splitmix(uint64_t& x) {
    uint64_t z = (x += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
  }
  explicit Rng(uint64_t seed) { for (auto& v : s) v = splitmix(seed); }
  static uint64_t rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
  uint64_t next() {
    const uint64_t r = rotl(s[1] * 5, 7) * 9, t = s[1] << 17;
    s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3]; s[2] ^= t; s[3] = rotl(s[3], 45);
    return r;
  }
  double uniform() { return ((next() >> 11) + 0.5) * (1.0 / 9007199254740992.0); }  // (0,1)
};

inline uint64_t mix(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t x = a * 0x9E3779B97F4A7C15ull ^ (b + 0x632BE59BD9B4E019ull) * 0xC2B2AE3D27D4EB4Full ^ (c + 0x165667B19E3779F9ull);
  return Rng::splitmix(x);
}

struct Segment {
  uint32_t max_doc = 0;
  uint64_t total_num_tokens = 0;
  std::vector<uint32_t> lengths;      // doc length before encoding (kept for tf clipping)
  std::vector<uint8_t> fieldnorm_ids;
  std::unique_ptr<tq::FieldPostingsWriter> writer;
  std::vector<tq::TermInfoOut> terms;  // one per requested density, in request order
  std::vector<uint8_t> positions;     // record_option 2: the field's `.pos` body (every term's position stream, in term order)
  std::vector<std::pair<uint64_t, uint64_t>> pos_ranges;  // ... and every term's positions_range in it
};

}  // namespace

struct tqs_index {
  std::vector<Segment> segs;
  std::vector<double> densities;
};

template <class F>
static void parallel_for(size_t n, int n_threads, F f) {
  if (n_threads <= 1 || n <= 1) { for (size_t i = 0; i < n; ++i) f(i); return; }
  std::atomic<size_t> next{0};
  std::vector<std::thread> th;
  for (int t = 0; t < n_threads; ++t) th.emplace_back([&] { for (;;) { size_t i = next.fetch_add(1); if (i >= n) break; f(i); } });
  for (auto& t : th) t.join();
}

extern "C" {

// densities[i] in (0, 1]: fraction of the segment's docs containing term i.
// record_option: 1 (WithFreqs) or 2 (WithFreqsAndPositions: every posting also gets tf distinct positions, uniform over the doc's
// length, written like PositionSerializer does -- src/positions/serializer.rs; tqs_positions / tqs_term_pos).
// segment_base: global ordinal of the first generated segment (seeds depend on the global ordinal, so a
// rank that generates only its own shard gets the same bytes as a full generation would give it).
tqs_index* tqs_generate(uint32_t n_segments, uint32_t docs_per_segment, const double* densities, uint32_t n_terms,
                        uint64_t seed, int record_option, int n_threads, uint32_t segment_base, uint32_t segment_stride) {
  auto* ix = new tqs_index();
  ix->segs.resize(n_segments);
  ix->densities.assign(densities, densities + n_terms);
  const double mu = std::log(80.0), sigma = 0.6;
  constexpr uint32_t CHUNK = 1u << 16;
  // 1. doc lengths / fieldnorm ids
  for (uint32_t s = 0; s < n_segments; ++s) {
    Segment& sg = ix->segs[s];
    sg.max_doc = docs_per_segment;
    sg.lengths.resize(docs_per_segment);
    sg.fieldnorm_ids.resize(docs_per_segment);
    const size_t n_chunks = (docs_per_segment + CHUNK - 1) / CHUNK;
    std::vector<uint64_t> partial(n_chunks, 0);
    parallel_for(n_chunks, n_threads, [&](size_t c) {
      Rng rng(mix(seed + segment_base + (uint64_t)s * segment_stride, 0xF1E1D0, c));
      const uint32_t lo = (uint32_t)c * CHUNK, hi = std::min<uint64_t>((uint64_t)lo + CHUNK, docs_per_segment);
      uint64_t sum = 0;
      for (uint32_t d = lo; d < hi; d += 2) {
        const double u1 = rng.uniform(), u2 = rng.uniform();
        const double r = std::sqrt(-2.0 * std::log(u1)), a = 6.283185307179586 * u2;
        const double z[2] = {r * std::cos(a), r * std::sin(a)};
        for (uint32_t k = 0; k < 2 && d + k < hi; ++k) {
          double len = std::floor(std::exp(mu + sigma * z[k]) + 0.5);
          if (len < 1.0) len = 1.0;
          if (len > 4096.0) len = 4096.0;
          sg.lengths[d + k] = (uint32_t)len;
          sg.fieldnorm_ids[d + k] = tq::fieldnorm_to_id((uint32_t)len);
          sum += (uint32_t)len;
        }
      }
      partial[c] = sum;
    });
    for (uint64_t p : partial) sg.total_num_tokens += p;
    sg.writer.reset(new tq::FieldPostingsWriter(record_option, sg.total_num_tokens, sg.fieldnorm_ids.data(), sg.max_doc));
    sg.terms.resize(n_terms);
  }
  // 2. posting lists: one task per (segment, term), encoded independently then appended in order
  const size_t n_tasks = (size_t)n_segments * n_terms;
  std::vector<std::vector<uint8_t>> encoded(n_tasks);
  std::vector<std::vector<uint8_t>> encoded_pos(record_option == 2 ? n_tasks : 0);
  std::vector<uint32_t> doc_freqs(n_tasks, 0);
  parallel_for(n_tasks, n_threads, [&](size_t task) {
    const uint32_t s = (uint32_t)(task / n_terms), t = (uint32_t)(task % n_terms);
    Segment& sg = ix->segs[s];
    const double p = std::min(1.0, std::max(1e-12, densities[t]));
    Rng rng(mix(seed + segment_base + (uint64_t)s * segment_stride, 0x7E63, t));
    std::vector<uint32_t> docs, tfs, deltas;
    docs.reserve((size_t)(p * sg.max_doc * 1.05) + 16);
    tfs.reserve(docs.capacity());
    if (record_option == 2) deltas.reserve(docs.capacity() * 2);
    const double inv_log_q = p < 1.0 ? 1.0 / std::log1p(-p) : 0.0;
    const double inv_log_tf = 1.0 / std::log(0.3);  // geometric(p=0.7): P(extra >= j) = 0.3^j
    uint64_t doc = 0;
    bool first = true;
    for (;;) {
      uint64_t gap = 1;
      if (p < 1.0) gap = 1 + (uint64_t)std::floor(std::log(rng.uniform()) * inv_log_q);
      doc = first ? gap - 1 : doc + gap;
      first = false;
      if (doc >= sg.max_doc) break;
      uint32_t tf = 1 + (uint32_t)std::floor(std::log(rng.uniform()) * inv_log_tf);
      if (tf > 10) tf = 10;
      if (tf > sg.lengths[doc]) tf = sg.lengths[doc];
      docs.push_back((uint32_t)doc);
      tfs.push_back(tf);
      if (record_option == 2) {
        // tf distinct positions in [0, length): sorted draws, made distinct by insertion; stored as first position + gaps
        uint32_t pos[10];
        const uint32_t len = sg.lengths[doc];
        for (uint32_t i = 0; i < tf; ++i) {
          uint32_t v = (uint32_t)(rng.uniform() * (len - i));  // i-th draw among the len - i free positions
          uint32_t j = 0;
          while (j < i && pos[j] <= v) { ++v; ++j; }
          for (uint32_t m = i; m > j; --m) pos[m] = pos[m - 1];
          pos[j] = v;
        }
        for (uint32_t i = 0; i < tf; ++i) deltas.push_back(i ? pos[i] - pos[i - 1] : pos[0]);
      }
    }
    doc_freqs[task] = (uint32_t)docs.size();
    sg.writer->encode_term(docs.data(), tfs.data(), (uint32_t)docs.size(), encoded[task]);
    if (record_option == 2) tq::encode_positions(deltas.data(), deltas.size(), encoded_pos[task]);
  });
  for (uint32_t s = 0; s < n_segments; ++s) {
    Segment& sg = ix->segs[s];
    size_t total = 0;
    for (uint32_t t = 0; t < n_terms; ++t) total += encoded[(size_t)s * n_terms + t].size();
    sg.writer->body_mut().reserve(8 + total + 64);
    for (uint32_t t = 0; t < n_terms; ++t) {
      const size_t task = (size_t)s * n_terms + t;
      sg.terms[t] = sg.writer->add_encoded(encoded[task], doc_freqs[task]);
      std::vector<uint8_t>().swap(encoded[task]);
      if (record_option == 2) {
        sg.pos_ranges.push_back({sg.positions.size(), sg.positions.size() + encoded_pos[task].size()});
        sg.positions.insert(sg.positions.end(), encoded_pos[task].begin(), encoded_pos[task].end());
        std::vector<uint8_t>().swap(encoded_pos[task]);
      }
    }
    std::vector<uint32_t>().swap(sg.lengths);
  }
  return ix;
}

void tqs_destroy(tqs_index* ix) { delete ix; }
uint32_t tqs_num_segments(tqs_index* ix) { return (uint32_t)ix->segs.size(); }
uint32_t tqs_max_doc(tqs_index* ix, uint32_t s) { return ix->segs[s].max_doc; }
uint64_t tqs_total_num_tokens(tqs_index* ix, uint32_t s) { return ix->segs[s].total_num_tokens; }
void tqs_body(tqs_index* ix, uint32_t s, const uint8_t** p, size_t* len) { *p = ix->segs[s].writer->body().data(); *len = ix->segs[s].writer->body().size(); }
void tqs_fieldnorm(tqs_index* ix, uint32_t s, const uint8_t** p, size_t* len) { *p = ix->segs[s].fieldnorm_ids.data(); *len = ix->segs[s].fieldnorm_ids.size(); }
void tqs_positions(tqs_index* ix, uint32_t s, const uint8_t** p, size_t* len) { *p = ix->segs[s].positions.data(); *len = ix->segs[s].positions.size(); }
void tqs_term_pos(tqs_index* ix, uint32_t s, uint32_t term, uint64_t* start, uint64_t* end) {
  const Segment& sg = ix->segs[s];
  if (term < sg.pos_ranges.size()) { *start = sg.pos_ranges[term].first; *end = sg.pos_ranges[term].second; } else { *start = *end = 0; }
}
void tqs_term_info(tqs_index* ix, uint32_t s, uint32_t term, uint32_t* doc_freq, uint64_t* start, uint64_t* end) {
  const tq::TermInfoOut& ti = ix->segs[s].terms[term];
  *doc_freq = ti.doc_freq; *start = ti.postings_start; *end = ti.postings_end;
}

}  // extern "C"
