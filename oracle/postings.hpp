// ORACLE — TEST INFRASTRUCTURE ONLY (see codec.hpp header).
// CPU restatement of the reference's BM25, fieldnorm code, skip list, postings serializer and
// posting cursors.  Follows (relative to /root/reference):
//   src/fieldnorm/code.rs:2-13            id_to_fieldnorm / fieldnorm_to_id / table
//   src/query/bm25.rs:7-193               K1, B, idf, tf cache, Bm25Weight
//   src/postings/skip.rs:9-302            bit-width byte, SkipSerializer, SkipReader, BlockInfo
//   src/postings/serializer.rs:303-481    Block, PostingsSerializer
//   src/postings/block_segment_postings.rs:12-420   BlockSegmentPostings
//   src/postings/segment_postings.rs:160-230        SegmentPostings (DocSet + Postings)
#pragma once
#include <algorithm>
#include <cassert>
#include <cmath>
#include <memory>
#include <stdexcept>
#include <vector>

#include "codec.hpp"

namespace tqo {

typedef float Score;
enum IndexRecordOption { Basic = 0, WithFreqs = 1, WithFreqsAndPositions = 2 };
inline bool has_freq(IndexRecordOption m) { return m != Basic; }
inline bool has_positions(IndexRecordOption m) { return m == WithFreqsAndPositions; }

// ---- fieldnorm code (Lucene SmallFloat byte4; verified against the reference's 256-entry
// table by tests/golden/gen_golden.py) --------------------------------------------------------
inline uint32_t id_to_fieldnorm(uint8_t id) {
  if (id < 24) return id;
  const uint32_t j = id - 24u, bits = j & 7u;
  const int shift = (int)(j >> 3) - 1;
  if (shift < 0) return 24u + bits;
  return 24u + ((bits | 8u) << shift);
}
inline uint8_t fieldnorm_to_id(uint32_t fieldnorm) {
  // binary_search(..).unwrap_or_else(|idx| idx - 1): index of the last table entry <= fieldnorm
  int lo = 0, hi = 255;
  while (lo < hi) {
    int mid = (lo + hi + 1) >> 1;
    if (id_to_fieldnorm((uint8_t)mid) <= fieldnorm) lo = mid; else hi = mid - 1;
  }
  return (uint8_t)lo;
}

struct FieldNormReader {
  const uint8_t* data = nullptr;  // one id per doc; nullptr => constant
  uint32_t num_docs_ = 0;
  uint8_t const_id = 0;
  static FieldNormReader from_data(const uint8_t* d, uint32_t n) { FieldNormReader r; r.data = d; r.num_docs_ = n; return r; }
  static FieldNormReader constant(uint32_t num_docs, uint32_t fieldnorm) {
    FieldNormReader r; r.num_docs_ = num_docs; r.const_id = fieldnorm_to_id(fieldnorm); return r;
  }
  uint32_t num_docs() const { return num_docs_; }
  uint8_t fieldnorm_id(uint32_t doc) const { return data ? data[doc] : const_id; }
};

// ---- BM25 (src/query/bm25.rs) ----------------------------------------------------------------
constexpr Score K1 = 1.2f;
constexpr Score B = 0.75f;

inline Score idf(uint64_t doc_freq, uint64_t doc_count) {
  if (doc_count < doc_freq) throw std::invalid_argument("doc_count >= doc_freq");
  const Score x = ((Score)(doc_count - doc_freq) + 0.5f) / ((Score)doc_freq + 0.5f);
  return std::log(1.0f + x);  // f32::ln
}
inline Score cached_tf_component(uint32_t fieldnorm, Score average_fieldnorm) {
  return K1 * (1.0f - B + B * (Score)fieldnorm / average_fieldnorm);
}
struct Bm25Weight {
  Score weight = 0;
  Score cache[256];
  Score average_fieldnorm = 0;
  static Bm25Weight from_idf(Score idf_value, Score average_fieldnorm) {
    Bm25Weight w;
    w.weight = idf_value * (1.0f + K1);
    w.average_fieldnorm = average_fieldnorm;
    for (int id = 0; id < 256; ++id) w.cache[id] = cached_tf_component(id_to_fieldnorm((uint8_t)id), average_fieldnorm);
    return w;
  }
  static Bm25Weight for_one_term(uint64_t term_doc_freq, uint64_t total_num_docs, Score avg_fieldnorm) {
    return from_idf(idf(term_doc_freq, total_num_docs), avg_fieldnorm);
  }
  Bm25Weight boost_by(Score boost) const {
    if (boost == 1.0f) return *this;
    Bm25Weight w = *this; w.weight = weight * boost; return w;
  }
  inline Score tf_factor(uint8_t fieldnorm_id, uint32_t term_freq) const {
    const Score tf = (Score)term_freq;
    const Score norm = cache[fieldnorm_id];
    return tf / (tf + norm);
  }
  inline Score score(uint8_t fieldnorm_id, uint32_t term_freq) const { return weight * tf_factor(fieldnorm_id, term_freq); }
  Score max_score() const { return score(255u, 2013265944u); }
};

// ---- skip list (src/postings/skip.rs) ---------------------------------------------------------
inline uint8_t encode_bitwidth(uint8_t bitwidth, bool delta_1) {
  assert(bitwidth < 32);
  return (uint8_t)(bitwidth | ((delta_1 ? 1 : 0) << 6));
}
inline void decode_bitwidth(uint8_t raw, uint8_t* bitwidth, bool* delta_1) {
  *delta_1 = ((raw >> 6) & 1) != 0;
  *bitwidth = raw & 0x1f;
}
inline uint8_t encode_block_wand_max_tf(uint32_t max_tf) { return (uint8_t)std::min<uint32_t>(max_tf, 255u); }
inline uint32_t decode_block_wand_max_tf(uint8_t code) { return code == 255 ? 0xFFFFFFFFu : (uint32_t)code; }
inline uint32_t read_u32(const uint8_t* p) { uint32_t v; std::memcpy(&v, p, 4); return v; }
inline void write_u32(uint32_t v, std::vector<uint8_t>& buf) { uint8_t b[4]; std::memcpy(b, &v, 4); buf.insert(buf.end(), b, b + 4); }

struct SkipSerializer {
  std::vector<uint8_t> buffer;
  void write_doc(uint32_t last_doc, uint8_t doc_num_bits) { write_u32(last_doc, buffer); buffer.push_back(encode_bitwidth(doc_num_bits, true)); }
  void write_term_freq(uint8_t tf_num_bits) { buffer.push_back(tf_num_bits); }
  void write_total_term_freq(uint32_t tf_sum) { write_u32(tf_sum, buffer); }
  void write_blockwand_max(uint8_t fieldnorm_id, uint32_t term_freq) { buffer.push_back(fieldnorm_id); buffer.push_back(encode_block_wand_max_tf(term_freq)); }
  void clear() { buffer.clear(); }
};

struct BlockInfo {
  bool bitpacked = false;
  uint8_t doc_num_bits = 0; bool strict_delta_encoded = false; uint8_t tf_num_bits = 0; uint32_t tf_sum = 0;
  uint8_t block_wand_fieldnorm_id = 0; uint32_t block_wand_term_freq = 0;
  uint32_t num_docs = 0;  // VInt variant
  static BlockInfo vint(uint32_t n) { BlockInfo b; b.bitpacked = false; b.num_docs = n; return b; }
};

struct SkipReader {
  uint32_t last_doc_in_block_ = 0;
  uint32_t last_doc_in_previous_block = 0;
  const uint8_t* read = nullptr;  // owned_read cursor
  IndexRecordOption skip_info = Basic;
  size_t byte_offset_ = 0;
  uint32_t remaining_docs_ = 0;
  BlockInfo block_info_;
  uint64_t position_offset_ = 0;

  SkipReader() {}
  SkipReader(const uint8_t* data, uint32_t doc_freq, IndexRecordOption info) { reset_all(data, doc_freq, info); }
  void reset_all(const uint8_t* data, uint32_t doc_freq, IndexRecordOption info) {
    last_doc_in_block_ = doc_freq >= COMPRESSION_BLOCK_SIZE ? 0 : TERMINATED;
    last_doc_in_previous_block = 0;
    read = data; skip_info = info;
    block_info_ = BlockInfo::vint(doc_freq);
    byte_offset_ = 0; remaining_docs_ = doc_freq; position_offset_ = 0;
    if (doc_freq >= COMPRESSION_BLOCK_SIZE) read_block_info();
  }
  bool has_remaining_docs() const { return remaining_docs_ != 0; }
  bool block_max_score(const Bm25Weight& w, Score* out) const {
    if (!block_info_.bitpacked) return false;
    *out = w.score(block_info_.block_wand_fieldnorm_id, block_info_.block_wand_term_freq);
    return true;
  }
  uint32_t last_doc_in_block() const { return last_doc_in_block_; }
  uint32_t remaining_docs() const { return remaining_docs_; }
  size_t byte_offset() const { return byte_offset_; }
  const BlockInfo& block_info() const { return block_info_; }

  void read_block_info() {
    const uint8_t* bytes = read;
    size_t advance_len;
    last_doc_in_block_ = read_u32(bytes);
    BlockInfo bi; bi.bitpacked = true;
    decode_bitwidth(bytes[4], &bi.doc_num_bits, &bi.strict_delta_encoded);
    switch (skip_info) {
      case Basic: advance_len = 5; break;
      case WithFreqs:
        bi.tf_num_bits = bytes[5]; bi.block_wand_fieldnorm_id = bytes[6];
        bi.block_wand_term_freq = decode_block_wand_max_tf(bytes[7]); advance_len = 8; break;
      default:
        bi.tf_num_bits = bytes[5]; bi.tf_sum = read_u32(bytes + 6); bi.block_wand_fieldnorm_id = bytes[10];
        bi.block_wand_term_freq = decode_block_wand_max_tf(bytes[11]); advance_len = 12; break;
    }
    block_info_ = bi;
    read += advance_len;
  }
  bool seek(uint32_t target) {
    if (last_doc_in_block() >= target) return false;
    for (;;) { advance(); if (last_doc_in_block() >= target) return true; }
  }
  void advance() {
    if (block_info_.bitpacked) {
      remaining_docs_ -= COMPRESSION_BLOCK_SIZE;
      byte_offset_ += compressed_block_size((uint8_t)(block_info_.doc_num_bits + block_info_.tf_num_bits));
      position_offset_ += block_info_.tf_sum;
    } else {
      remaining_docs_ = 0;
      byte_offset_ = (size_t)-1;
    }
    last_doc_in_previous_block = last_doc_in_block_;
    if (remaining_docs_ >= COMPRESSION_BLOCK_SIZE) read_block_info();
    else { last_doc_in_block_ = TERMINATED; block_info_ = BlockInfo::vint(remaining_docs_); }
  }
};

// ---- serializer (src/postings/serializer.rs:303-481) ----------------------------------------
struct PostingsSerializer {
  uint32_t last_doc_id_encoded = 0;
  BlockEncoder block_encoder;
  uint32_t blk_docs[128], blk_tfs[128]; size_t blk_len = 0;
  std::vector<uint8_t> postings_write;
  SkipSerializer skip_write;
  IndexRecordOption mode;
  bool have_fieldnorms; FieldNormReader fieldnorm_reader;
  bool have_bm25 = false; Bm25Weight bm25_weight;
  Score avg_fieldnorm;
  bool term_has_freq = false;

  PostingsSerializer(Score avg, IndexRecordOption m, const FieldNormReader* fnr)
      : mode(m), have_fieldnorms(fnr != nullptr), avg_fieldnorm(avg) { if (fnr) fieldnorm_reader = *fnr; }

  void new_term(uint32_t term_doc_freq, bool record_term_freq) {
    // FieldSerializer::new_term calls postings_serializer.clear() first (serializer.rs:194)
    blk_len = 0; last_doc_id_encoded = 0;
    have_bm25 = false;
    term_has_freq = has_freq(mode) && record_term_freq;
    if (!term_has_freq) return;
    if (!have_fieldnorms) return;
    const uint64_t num_docs_in_segment = fieldnorm_reader.num_docs();
    if (num_docs_in_segment == 0) return;
    bm25_weight = Bm25Weight::for_one_term(term_doc_freq, num_docs_in_segment, avg_fieldnorm);
    have_bm25 = true;
  }
  void write_block() {
    {
      size_t written;
      uint8_t num_bits = block_encoder.compress_block_sorted(blk_docs, last_doc_id_encoded, &written);
      last_doc_id_encoded = blk_docs[127];
      skip_write.write_doc(last_doc_id_encoded, num_bits);
      postings_write.insert(postings_write.end(), block_encoder.output, block_encoder.output + written);
    }
    if (term_has_freq) {
      size_t written;
      uint8_t num_bits = block_encoder.compress_block_unsorted(blk_tfs, true, &written);
      postings_write.insert(postings_write.end(), block_encoder.output, block_encoder.output + written);
      skip_write.write_term_freq(num_bits);
      if (has_positions(mode)) {
        uint32_t sum_freq = 0;
        for (int i = 0; i < 128; ++i) sum_freq += blk_tfs[i];
        skip_write.write_total_term_freq(sum_freq);
      }
      uint8_t bw_fn = 0; uint32_t bw_tf = 0;
      if (have_bm25 && have_fieldnorms) {
        // Iterator::max_by keeps the LAST of several equal maxima; partial_cmp -> Equal on NaN.
        Score best = 0; bool first = true;
        for (int i = 0; i < 128; ++i) {
          const uint8_t fid = fieldnorm_reader.fieldnorm_id(blk_docs[i]);
          const Score s = bm25_weight.tf_factor(fid, blk_tfs[i]);
          const bool greater = (!first) && (best > s);  // Ordering::Greater keeps the old one
          if (first || !greater) { best = s; bw_fn = fid; bw_tf = blk_tfs[i]; }
          first = false;
        }
      }
      skip_write.write_blockwand_max(bw_fn, bw_tf);
    }
    blk_len = 0;
  }
  void write_doc(uint32_t doc_id, uint32_t term_freq) {
    blk_docs[blk_len] = doc_id; blk_tfs[blk_len] = term_freq; ++blk_len;
    if (blk_len == COMPRESSION_BLOCK_SIZE) write_block();
  }
  void close_term(uint32_t doc_freq, std::vector<uint8_t>& output_write) {
    if (blk_len != 0) {
      size_t n = block_encoder.compress_vint_sorted(blk_docs, blk_len, last_doc_id_encoded);
      postings_write.insert(postings_write.end(), block_encoder.output, block_encoder.output + n);
      if (term_has_freq) {
        n = block_encoder.compress_vint_unsorted(blk_tfs, blk_len);
        postings_write.insert(postings_write.end(), block_encoder.output, block_encoder.output + n);
      }
      blk_len = 0;
    }
    if (doc_freq >= COMPRESSION_BLOCK_SIZE) {
      common_vint_serialize(skip_write.buffer.size(), output_write);
      output_write.insert(output_write.end(), skip_write.buffer.begin(), skip_write.buffer.end());
    }
    output_write.insert(output_write.end(), postings_write.begin(), postings_write.end());
    skip_write.clear(); postings_write.clear(); have_bm25 = false;
  }
};

// ---- BlockSegmentPostings (src/postings/block_segment_postings.rs) ---------------------------
enum FreqReadingOption { NoFreq, SkipFreq, ReadFreq };

struct BlockSegmentPostings {
  BlockDecoder doc_decoder{TERMINATED};
  bool block_loaded = false;
  BlockDecoder freq_decoder{1};
  FreqReadingOption freq_reading_option = NoFreq;
  bool block_max_cached = false; Score block_max_score_cache = 0;
  uint32_t doc_freq_ = 0;
  const uint8_t* data = nullptr; size_t data_len = 0;
  SkipReader skip_reader;

  static BlockSegmentPostings empty() {
    BlockSegmentPostings p; p.block_loaded = true; p.skip_reader = SkipReader(nullptr, 0, Basic); return p;
  }
  // open(doc_freq, data, record_option, requested_option)   block_segment_postings.rs:97-140
  static BlockSegmentPostings open(uint32_t doc_freq, const uint8_t* bytes, size_t len,
                                   IndexRecordOption record_option, IndexRecordOption requested_option) {
    BlockSegmentPostings p;
    const uint8_t* skip_data = nullptr; size_t skip_len = 0;
    const uint8_t* postings_data = bytes; size_t postings_len = len;
    if (doc_freq >= COMPRESSION_BLOCK_SIZE) {  // split_into_skips_and_postings
      uint64_t sl; size_t consumed;
      if (!common_vint_deserialize(bytes, len, &sl, &consumed)) throw std::runtime_error("corrupt skip_len");
      skip_len = (size_t)sl; skip_data = bytes + consumed;
      postings_data = skip_data + skip_len; postings_len = len - consumed - skip_len;
      const size_t block_count = doc_freq / COMPRESSION_BLOCK_SIZE;
      if (skip_len < 8 * block_count) record_option = Basic;  // :116-123
      p.skip_reader = SkipReader(skip_data, doc_freq, record_option);
    } else {
      p.skip_reader = SkipReader(nullptr, doc_freq, record_option);
    }
    if (record_option == Basic) p.freq_reading_option = NoFreq;
    else if (requested_option == Basic) p.freq_reading_option = SkipFreq;
    else p.freq_reading_option = ReadFreq;
    p.doc_freq_ = doc_freq; p.data = postings_data; p.data_len = postings_len;
    p.load_block();
    return p;
  }
  uint32_t doc_freq() const { return doc_freq_; }
  const uint32_t* docs() const { return doc_decoder.output; }
  uint32_t doc(size_t idx) const { return doc_decoder.output[idx]; }
  uint32_t freq(size_t idx) const { return freq_decoder.output[idx]; }
  size_t block_len() const { return doc_decoder.output_len; }
  bool block_is_loaded() const { return block_loaded; }
  bool has_remaining_docs() const { return skip_reader.has_remaining_docs(); }

  Score block_max_score(const FieldNormReader& fnr, const Bm25Weight& w) {  // :147-179
    if (block_max_cached) return block_max_score_cache;
    Score s;
    if (skip_reader.block_max_score(w, &s)) { block_max_cached = true; block_max_score_cache = s; return s; }
    if (block_is_loaded()) {
      bool any = false; Score best = 0.0f;
      for (size_t i = 0; i < doc_decoder.output_len; ++i) {
        const Score sc = w.score(fnr.fieldnorm_id(doc_decoder.output[i]), freq_decoder.output[i]);
        best = any ? std::fmax(best, sc) : sc;  // fold(first, Score::max)
        any = true;
      }
      block_max_cached = true; block_max_score_cache = any ? best : 0.0f;
      return block_max_score_cache;
    }
    return w.max_score();
  }
  size_t seek(uint32_t target_doc) {  // :271-288
    seek_block(target_doc);
    load_block();
    return search_block(doc_decoder.output, target_doc);
  }
  void seek_block(uint32_t target_doc) {  // :327-332
    if (skip_reader.seek(target_doc)) { block_max_cached = false; block_loaded = false; }
  }
  void load_block() {  // :343-391
    if (block_loaded) return;
    const size_t offset = skip_reader.byte_offset();
    const BlockInfo& bi = skip_reader.block_info();
    const bool read_freq = (freq_reading_option == ReadFreq);
    if (bi.bitpacked) {
      const uint8_t* d = data + offset;
      const size_t consumed = doc_decoder.uncompress_block_sorted(d, skip_reader.last_doc_in_previous_block, bi.doc_num_bits, bi.strict_delta_encoded);
      if (read_freq) freq_decoder.uncompress_block_unsorted(d + consumed, bi.tf_num_bits, bi.strict_delta_encoded);
    } else {
      const uint8_t* d = nullptr; size_t dlen = 0;
      if (bi.num_docs != 0) { d = data + offset; dlen = data_len - offset; }
      const size_t consumed = doc_decoder.uncompress_vint_sorted(d, skip_reader.last_doc_in_previous_block, bi.num_docs, TERMINATED);
      if (read_freq && dlen > consumed) freq_decoder.uncompress_vint_unsorted(d + consumed, bi.num_docs, TERMINATED);
    }
    block_loaded = true;
  }
  void advance() {  // :394-399
    skip_reader.advance();
    block_loaded = false; block_max_cached = false;
    load_block();
  }
};

// ---- SegmentPostings (src/postings/segment_postings.rs:160-230) -----------------------------
struct SegmentPostings {
  BlockSegmentPostings block_cursor;
  size_t cur = 0;
  uint32_t doc() const { return block_cursor.doc(cur); }
  uint32_t advance() {
    if (cur == COMPRESSION_BLOCK_SIZE - 1) { cur = 0; block_cursor.advance(); }
    else cur += 1;
    return doc();
  }
  uint32_t seek(uint32_t target) {
    if (doc() >= target) return doc();
    cur = std::min<size_t>(cur + 1, COMPRESSION_BLOCK_SIZE - 1);
    if (doc() >= target) return doc();
    cur = block_cursor.seek(target);
    return doc();
  }
  uint32_t size_hint() const { return block_cursor.doc_freq(); }
  uint32_t term_freq() const { return block_cursor.freq(cur); }
};

}  // namespace tqo
