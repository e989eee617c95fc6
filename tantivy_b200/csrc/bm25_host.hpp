// Host-side BM25 scalars and the fieldnorm code of the product.
// Same arithmetic, same operation order and same f32 type as the reference:
//   src/query/bm25.rs:7-8,52-69,141-175   K1, B, idf, cached_tf_component, weight, tf_factor
//   src/fieldnorm/code.rs:2-13            id <-> fieldnorm (Lucene SmallFloat byte4 form)
// Compiled with -ffp-contract=off / -fmad=false so that no a*b+c is fused.
#pragma once
#include <cmath>
#include <cstdint>

#ifdef __CUDACC__
#define TQ_HD __host__ __device__
#else
#define TQ_HD
#endif

namespace tq {

constexpr float BM25_K1 = 1.2f;
constexpr float BM25_B = 0.75f;

// 24 exact small values, then 3-bit mantissa / 5-bit exponent steps.
TQ_HD inline uint32_t id_to_fieldnorm(uint32_t id) {
  if (id < 24u) return id;
  const uint32_t j = id - 24u;
  const uint32_t mant = j & 7u;
  const uint32_t e = j >> 3;
  return e == 0 ? 24u + mant : 24u + ((mant | 8u) << (e - 1u));
}

inline uint8_t fieldnorm_to_id(uint32_t fieldnorm) {
  // largest id whose fieldnorm is <= the requested one
  uint32_t lo = 0, hi = 256;
  while (hi - lo > 1) {
    const uint32_t mid = (lo + hi) / 2;
    if (id_to_fieldnorm(mid) <= fieldnorm) lo = mid; else hi = mid;
  }
  return (uint8_t)lo;
}

inline float bm25_idf(uint64_t doc_freq, uint64_t doc_count) {
  const float x = ((float)(doc_count - doc_freq) + 0.5f) / ((float)doc_freq + 0.5f);
  return logf(1.0f + x);  // host logf, never a device approximation
}

inline float bm25_weight(uint64_t doc_freq, uint64_t doc_count, float boost) {
  float w = bm25_idf(doc_freq, doc_count) * (1.0f + BM25_K1);
  if (boost != 1.0f) w = w * boost;
  return w;
}

inline float bm25_tf_norm(uint32_t fieldnorm, float average_fieldnorm) {
  return BM25_K1 * (1.0f - BM25_B + BM25_B * (float)fieldnorm / average_fieldnorm);
}

inline void bm25_tf_cache(float average_fieldnorm, float out[256]) {
  for (uint32_t id = 0; id < 256; ++id) out[id] = bm25_tf_norm(id_to_fieldnorm(id), average_fieldnorm);
}

}  // namespace tq
