// Host-side writer of tantivy-format posting lists (product code; used to build the synthetic
// benchmark segments and test segments that the device path consumes).
//
// Produces, byte for byte, what the reference's write side produces:
//   src/postings/serializer.rs:120-133   field body = u64 LE total_num_tokens, then the terms
//   src/postings/serializer.rs:353-481   per term: [VInt(skip_len) skip records] blocks [VInt tail]
//   src/postings/skip.rs:55-90           skip record: u32 last_doc, u8 doc_bits|0x40, [u8 tf_bits],
//                                        [u32 tf_sum], [u8 blockwand fieldnorm id, u8 blockwand tf]
//   src/postings/compression/mod.rs:36-76  strict-delta doc blocks, (tf-1) blocks
//   crate bitpacking 0.9 BitPacker4x     4 interleaved bit streams, one 16-byte vector per word
//   src/postings/compression/vint.rs     tail: 7-bit groups, 0x80 marks the LAST byte
// This encoder is written independently of oracle/ (lane-streaming accumulators instead of bit
// position arithmetic); tests check both produce identical bytes.
#pragma once
#include <cstdint>
#include <cstring>
#include <vector>

#include "bm25_host.hpp"

namespace tq {

constexpr uint32_t kBlock = 128;

inline uint32_t bits_for(uint32_t or_of_values) { return or_of_values ? 32u - (uint32_t)__builtin_clz(or_of_values) : 0u; }

// Packs 128 values of `b` bits: value j goes to stream j%4; stream word w lands at out[4*w + stream].
inline void pack_block_4x(const uint32_t* vals, uint32_t b, uint8_t* out) {
  if (b == 0) return;
  uint32_t* o = reinterpret_cast<uint32_t*>(out);  // callers hand 4-byte aligned scratch
  for (uint32_t lane = 0; lane < 4; ++lane) {
    uint64_t acc = 0;
    uint32_t filled = 0, w = 0;
    for (uint32_t row = 0; row < 32; ++row) {
      acc |= (uint64_t)vals[4 * row + lane] << filled;
      filled += b;
      if (filled >= 32) {
        o[4 * w + lane] = (uint32_t)acc;
        acc >>= 32;
        filled -= 32;
        ++w;
      }
    }
  }
}

inline void put_vint(uint32_t v, std::vector<uint8_t>& out) {
  while (v >= 128u) { out.push_back((uint8_t)(v & 127u)); v >>= 7; }
  out.push_back((uint8_t)(v | 128u));
}
inline void put_vint64(uint64_t v, std::vector<uint8_t>& out) {
  while (v >= 128u) { out.push_back((uint8_t)(v & 127u)); v >>= 7; }
  out.push_back((uint8_t)(v | 128u));
}
inline void put_u32(uint32_t v, std::vector<uint8_t>& out) {
  out.push_back((uint8_t)v); out.push_back((uint8_t)(v >> 8)); out.push_back((uint8_t)(v >> 16)); out.push_back((uint8_t)(v >> 24));
}

// Writes one term's position deltas the way the reference's PositionSerializer does (src/positions/serializer.rs,
// layout src/positions/mod.rs:13-31): VInt(number of bit-packed blocks), their bit widths, the blocks of 128 deltas
// (plain values, not minus-one), then the rest as VInts.  `deltas` = for every posting, its first position followed by
// the gaps between consecutive positions.  Appends to `out`; the byte range appended is the term's positions_range.
inline void encode_positions(const uint32_t* deltas, size_t n, std::vector<uint8_t>& out) {
  const size_t n_blocks = n / kBlock;
  put_vint64(n_blocks, out);
  const size_t widths_at = out.size();
  out.resize(out.size() + n_blocks);
  alignas(16) uint8_t packed[kBlock * 4];
  for (size_t b = 0; b < n_blocks; ++b) {
    const uint32_t* d = deltas + b * kBlock;
    uint32_t orv = 0;
    for (uint32_t i = 0; i < kBlock; ++i) orv |= d[i];
    const uint32_t bits = bits_for(orv);
    out[widths_at + b] = (uint8_t)bits;
    pack_block_4x(d, bits, packed);
    out.insert(out.end(), packed, packed + 16 * bits);
  }
  for (size_t i = n_blocks * kBlock; i < n; ++i) put_vint(deltas[i], out);
}

struct TermInfoOut { uint32_t doc_freq; uint64_t postings_start, postings_end; };

// Encodes whole posting lists of one field of one segment.
class FieldPostingsWriter {
 public:
  // fieldnorm_ids may be null (no block-max information is then written: (0,0) pairs).
  FieldPostingsWriter(int record_option, uint64_t total_num_tokens, const uint8_t* fieldnorm_ids, uint32_t max_doc)
      : record_option_(record_option), fieldnorm_ids_(fieldnorm_ids), max_doc_(max_doc) {
    body_.resize(8);
    std::memcpy(body_.data(), &total_num_tokens, 8);
    if (fieldnorm_ids_ && max_doc_) {
      avg_fieldnorm_ = (float)total_num_tokens / (float)max_doc_;  // serializer.rs:130-133
      bm25_tf_cache(avg_fieldnorm_, norm_cache_);
    }
  }

  // Serializes one term into `out` (appended). docs strictly ascending; tfs >= 1 or null.
  // term_has_freq follows serializer.rs:374 (mode.has_freq() && record_term_freq).
  void encode_term(const uint32_t* docs, const uint32_t* tfs, uint32_t n, std::vector<uint8_t>& out) const {
    const bool with_freq = record_option_ != 0 && tfs != nullptr;
    const bool with_pos = record_option_ == 2;
    const bool with_blockmax = with_freq && fieldnorm_ids_ && max_doc_ != 0;
    const uint32_t n_blocks = n / kBlock;
    std::vector<uint8_t> skip, blocks;
    skip.reserve((size_t)n_blocks * 12);
    blocks.reserve((size_t)n * 2 + 64);
    alignas(16) uint32_t tmp[kBlock];
    alignas(16) uint8_t packed[kBlock * 4];
    uint32_t prev = 0;
    for (uint32_t b = 0; b < n_blocks; ++b) {
      const uint32_t* d = docs + (size_t)b * kBlock;
      // strictly-sorted deltas; before the very first doc the predecessor is "-1" (offset 0 => None)
      uint32_t last = (prev == 0) ? 0xFFFFFFFFu : prev, orv = 0;
      for (uint32_t i = 0; i < kBlock; ++i) { tmp[i] = d[i] - last - 1u; orv |= tmp[i]; last = d[i]; }
      const uint32_t db = bits_for(orv);
      pack_block_4x(tmp, db, packed);
      blocks.insert(blocks.end(), packed, packed + 16 * db);
      prev = d[kBlock - 1];
      put_u32(prev, skip);
      skip.push_back((uint8_t)(db | 0x40u));
      if (with_freq) {
        const uint32_t* f = tfs + (size_t)b * kBlock;
        uint32_t orf = 0, sum = 0;
        for (uint32_t i = 0; i < kBlock; ++i) { tmp[i] = f[i] - 1u; orf |= tmp[i]; sum += f[i]; }
        const uint32_t tb = bits_for(orf);
        pack_block_4x(tmp, tb, packed);
        blocks.insert(blocks.end(), packed, packed + 16 * tb);
        skip.push_back((uint8_t)tb);
        if (with_pos) put_u32(sum, skip);
        uint8_t bm_fn = 0; uint32_t bm_tf = 0;
        if (with_blockmax) {
          // argmax of tf/(tf+norm[fieldnorm_id]); on ties the later posting wins (Iterator::max_by)
          float best = 0.0f;
          for (uint32_t i = 0; i < kBlock; ++i) {
            const uint8_t id = fieldnorm_ids_[d[i]];
            const float tf = (float)f[i];
            const float s = tf / (tf + norm_cache_[id]);
            if (i == 0 || !(best > s)) { best = s; bm_fn = id; bm_tf = f[i]; }
          }
        }
        skip.push_back(bm_fn);
        skip.push_back((uint8_t)(bm_tf < 255u ? bm_tf : 255u));
      }
    }
    const uint32_t tail = n - n_blocks * kBlock;
    if (n >= kBlock) {
      put_vint64(skip.size(), out);
      out.insert(out.end(), skip.begin(), skip.end());
    }
    out.insert(out.end(), blocks.begin(), blocks.end());
    if (tail) {
      uint32_t last = prev;  // plain (non-strict) delta from last_doc_id_encoded
      for (uint32_t i = n - tail; i < n; ++i) { put_vint(docs[i] - last, out); last = docs[i]; }
      if (with_freq) for (uint32_t i = n - tail; i < n; ++i) put_vint(tfs[i], out);
    }
  }

  TermInfoOut add_term(const uint32_t* docs, const uint32_t* tfs, uint32_t n) {
    TermInfoOut ti;
    ti.doc_freq = n;
    ti.postings_start = body_.size() - 8;
    encode_term(docs, tfs, n, body_);
    ti.postings_end = body_.size() - 8;
    return ti;
  }
  // appends an already encoded term (used by the parallel generator)
  TermInfoOut add_encoded(const std::vector<uint8_t>& bytes, uint32_t doc_freq) {
    TermInfoOut ti;
    ti.doc_freq = doc_freq;
    ti.postings_start = body_.size() - 8;
    body_.insert(body_.end(), bytes.begin(), bytes.end());
    ti.postings_end = body_.size() - 8;
    return ti;
  }
  const std::vector<uint8_t>& body() const { return body_; }
  std::vector<uint8_t>& body_mut() { return body_; }

 private:
  int record_option_;
  const uint8_t* fieldnorm_ids_;
  uint32_t max_doc_;
  float avg_fieldnorm_ = 0.0f;
  float norm_cache_[256];
  std::vector<uint8_t> body_;
};

}  // namespace tq
