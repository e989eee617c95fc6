// ORACLE — TEST INFRASTRUCTURE ONLY (see codec.hpp header).
// CPU restatement of the reference's term-position codec, the data format of SURVEY.md §8(f) N3 (phrase queries):
//   src/positions/mod.rs:1-31        layout: VInt(#bit-packed blocks) | bit widths | bit-packed blocks of 128 position
//                                    deltas | VInt-encoded rest (read until the slice ends)
//   src/positions/serializer.rs      PositionSerializer::{write_positions_delta, flush_block, close_term}
//   src/positions/reader.rs          PositionReader::{open, advance_num_blocks, load_block, read}
// Deltas are stored as they come (unsorted, not minus-one: serializer.rs:66, reader.rs:94-95).
#pragma once
#include <algorithm>
#include <cstdint>
#include <vector>

#include "codec.hpp"

namespace tqo {

// src/postings/compression/vint.rs:85-107
inline size_t vint_uncompress_unsorted_until_end(const uint8_t* data, size_t len, uint32_t* output, size_t cap) {
  size_t read = 0;
  for (size_t i = 0; i < cap; ++i) {
    if (read == len) return i;
    uint32_t result = 0, shift = 0;
    for (;;) {
      const uint8_t b = data[read++];
      result += (uint32_t)(b % 128u) << shift;
      if (b & 128u) break;
      shift += 7;
    }
    output[i] = result;
  }
  return cap;
}

struct PositionSerializer {
  BlockEncoder block_encoder;
  std::vector<uint8_t> out;               // positions_wrt
  std::vector<uint8_t> positions_buffer;  // the current term's blocks
  std::vector<uint32_t> block;
  std::vector<uint8_t> bit_widths;

  uint64_t written_bytes() const { return out.size(); }

  void write_positions_delta(const uint32_t* deltas, size_t n) {
    while (n) {
      const size_t take = std::min<size_t>(COMPRESSION_BLOCK_SIZE - block.size(), n);
      block.insert(block.end(), deltas, deltas + take);
      deltas += take;
      n -= take;
      if (block.size() == COMPRESSION_BLOCK_SIZE) flush_block();
    }
  }
  void flush_block() {
    if (block.empty()) return;
    if (block.size() == COMPRESSION_BLOCK_SIZE) {
      size_t written = 0;
      const uint8_t bits = block_encoder.compress_block_unsorted(block.data(), false, &written);
      bit_widths.push_back(bits);
      positions_buffer.insert(positions_buffer.end(), block_encoder.output, block_encoder.output + written);
    } else {
      const size_t written = block_encoder.compress_vint_unsorted(block.data(), block.size());
      positions_buffer.insert(positions_buffer.end(), block_encoder.output, block_encoder.output + written);
    }
    block.clear();
  }
  void close_term() {
    flush_block();
    common_vint_serialize(bit_widths.size(), out);
    out.insert(out.end(), bit_widths.begin(), bit_widths.end());
    out.insert(out.end(), positions_buffer.begin(), positions_buffer.end());
    bit_widths.clear();
    positions_buffer.clear();
  }
};

struct PositionReader {
  const uint8_t* original_bit_widths = nullptr;
  size_t original_num_blocks = 0;
  const uint8_t* original_positions = nullptr;
  size_t original_positions_len = 0;
  const uint8_t* bit_widths = nullptr;  // consumed as the anchor advances
  size_t num_blocks = 0;
  const uint8_t* positions = nullptr;
  size_t positions_len = 0;
  BlockDecoder block_decoder;
  uint64_t block_offset = (uint64_t)INT64_MAX;
  uint64_t anchor_offset = 0;

  static bool open(const uint8_t* data, size_t len, PositionReader* r) {
    uint64_t n = 0;
    size_t used = 0;
    if (!common_vint_deserialize(data, len, &n, &used) || used + n > len) return false;
    r->original_bit_widths = data + used;
    r->original_num_blocks = (size_t)n;
    r->original_positions = data + used + n;
    r->original_positions_len = len - used - (size_t)n;
    r->reset();
    return true;
  }
  void reset() {
    bit_widths = original_bit_widths; num_blocks = original_num_blocks;
    positions = original_positions; positions_len = original_positions_len;
    block_offset = (uint64_t)INT64_MAX;
    anchor_offset = 0;
  }
  void advance_num_blocks(size_t n) {
    size_t bits = 0;
    for (size_t i = 0; i < n; ++i) bits += bit_widths[i];
    const size_t bytes = bits * COMPRESSION_BLOCK_SIZE / 8;
    bit_widths += n; num_blocks -= n;
    positions += bytes; positions_len -= bytes;
    anchor_offset += (uint64_t)n * COMPRESSION_BLOCK_SIZE;
  }
  void load_block(size_t block_rel_id) {
    size_t bits = 0;
    for (size_t i = 0; i < block_rel_id && i < num_blocks; ++i) bits += bit_widths[i];
    const size_t byte_offset = bits * COMPRESSION_BLOCK_SIZE / 8;
    const uint8_t* compressed = positions + byte_offset;
    if (num_blocks > block_rel_id) {
      block_decoder.uncompress_block_unsorted(compressed, bit_widths[block_rel_id], false);
    } else {
      block_decoder.output_len = vint_uncompress_unsorted_until_end(compressed, positions_len - byte_offset, block_decoder.output, COMPRESSION_BLOCK_SIZE);
    }
    block_offset = anchor_offset + (uint64_t)block_rel_id * COMPRESSION_BLOCK_SIZE;
  }
  // positions [offset, offset + n) of the term
  void read(uint64_t offset, uint32_t* output, size_t n) {
    if (offset < anchor_offset) reset();
    const int64_t delta_to_block = (int64_t)offset - (int64_t)block_offset;
    if (!(delta_to_block >= 0 && delta_to_block < 128)) {
      advance_num_blocks((size_t)((offset - anchor_offset) / COMPRESSION_BLOCK_SIZE));
      load_block(0);
    } else {
      advance_num_blocks((size_t)((block_offset - anchor_offset) / COMPRESSION_BLOCK_SIZE));
    }
    for (size_t i = 1;; ++i) {
      const size_t in_block = (size_t)(offset % COMPRESSION_BLOCK_SIZE);
      const size_t remaining = COMPRESSION_BLOCK_SIZE - in_block;
      if (remaining >= n) {
        std::copy(block_decoder.output + in_block, block_decoder.output + in_block + n, output);
        break;
      }
      std::copy(block_decoder.output + in_block, block_decoder.output + COMPRESSION_BLOCK_SIZE, output);
      output += remaining;
      n -= remaining;
      offset += remaining;
      load_block(i);
    }
  }
};

}  // namespace tqo
