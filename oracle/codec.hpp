// ORACLE — TEST INFRASTRUCTURE ONLY.  CPU restatement of the reference's posting codec.
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
// may use anything under oracle/.  The product path never links or calls this.
//
// Follows (paths relative to /root/reference):
//   src/postings/compression/mod.rs:1-168      BlockEncoder / BlockDecoder
//   src/postings/compression/vint.rs:1-127     postings VInt (stop bit 0x80 on the LAST byte)
//   common/src/vint.rs:59-112                  common VInt (same wire format) used for skip_len
//   external crate `bitpacking` ^0.9.3, BitPacker4x (Cargo.toml:42-44) — NOT in the tree.
//
// PARITY UNPINNED for the packed bytes of BitPacker4x: no test in the reference pins them
// (SURVEY.md §8c).  The layout restated here is the crate's published SIMD-BP128 "vertical"
// layout: 128 values = 32 rows x 4 lanes, value j -> lane j&3, row j>>2; every lane is an
// independent LSB-first bit stream of 32 b-bit values in b 32-bit words; output word index
// = w*4 + lane (one 16-byte vector per w), little endian; size = 16*b bytes
// (pinned in-tree: compression/mod.rs:13-15, skip.rs:284).
// Delta modes: sorted = v[j]-v[j-1] with v[-1]=initial; strictly sorted = v[j]-v[j-1]-1 with
// v[-1] = initial, None meaning u32::MAX (wrapping), selected when offset==0
// (compression/mod.rs:36-39,112-113).
#pragma once
#include <cstdint>
#include <cstring>
#include <cstddef>
#include <vector>
#include <emmintrin.h>

namespace tqo {

constexpr uint32_t COMPRESSION_BLOCK_SIZE = 128;  // BitPacker4x::BLOCK_LEN
constexpr uint32_t TERMINATED = 0x7FFFFFFFu;      // src/docset.rs:12

inline uint8_t bit_length(uint32_t v) { return v ? (uint8_t)(32 - __builtin_clz(v)) : 0; }

// compression/mod.rs:13-15
inline size_t compressed_block_size(uint8_t num_bits) { return (size_t)num_bits * COMPRESSION_BLOCK_SIZE / 8; }

// ---- BitPacker4x: definition-following scalar pack / unpack -------------------------------
inline void bp4x_pack(const uint32_t* v, uint8_t b, uint8_t* out) {
  uint32_t words[128];
  std::memset(words, 0, sizeof(uint32_t) * 4 * b);
  if (b == 0) return;
  for (uint32_t j = 0; j < 128; ++j) {
    const uint32_t lane = j & 3, row = j >> 2;
    const uint32_t bitpos = row * b, w = bitpos >> 5, sh = bitpos & 31;
    const uint32_t val = (b == 32) ? v[j] : (v[j] & ((1u << b) - 1));
    words[w * 4 + lane] |= val << sh;
    if (sh + b > 32) words[(w + 1) * 4 + lane] |= val >> (32 - sh);
  }
  std::memcpy(out, words, 16 * (size_t)b);
}

inline void bp4x_unpack_scalar(const uint8_t* in, uint8_t b, uint32_t* out) {
  if (b == 0) { std::memset(out, 0, 128 * sizeof(uint32_t)); return; }
  uint32_t words[132];
  std::memcpy(words, in, 16 * (size_t)b);
  words[4 * b] = words[4 * b + 1] = words[4 * b + 2] = words[4 * b + 3] = 0;
  const uint32_t mask = (b == 32) ? 0xFFFFFFFFu : ((1u << b) - 1);
  for (uint32_t j = 0; j < 128; ++j) {
    const uint32_t lane = j & 3, row = j >> 2;
    const uint32_t bitpos = row * b, w = bitpos >> 5, sh = bitpos & 31;
    uint64_t two = (uint64_t)words[w * 4 + lane] | ((uint64_t)words[(w + 1) * 4 + lane] << 32);
    out[j] = (uint32_t)(two >> sh) & mask;
  }
}

// SSE2 unpack, one 128-bit vector per row, as the crate's SSE3 back-end does; used by the
// timed CPU baseline.  Bit-identical to bp4x_unpack_scalar (checked in tests).
template <int B>
inline void bp4x_unpack_sse(const uint8_t* in, uint32_t* out) {
  if constexpr (B == 0) {
    std::memset(out, 0, 128 * sizeof(uint32_t));
  } else if constexpr (B == 32) {
    std::memcpy(out, in, 512);
  } else {
    const __m128i* src = reinterpret_cast<const __m128i*>(in);
    __m128i* dst = reinterpret_cast<__m128i*>(out);
    const __m128i mask = _mm_set1_epi32((int)((1u << B) - 1));
#pragma GCC unroll 32
    for (int row = 0; row < 32; ++row) {
      const int bitpos = row * B, w = bitpos >> 5, sh = bitpos & 31;
      __m128i v = _mm_srli_epi32(_mm_loadu_si128(src + w), sh);
      if (sh + B > 32) v = _mm_or_si128(v, _mm_slli_epi32(_mm_loadu_si128(src + w + 1), 32 - sh));
      _mm_storeu_si128(dst + row, _mm_and_si128(v, mask));
    }
  }
}

inline void bp4x_unpack(const uint8_t* in, uint8_t b, uint32_t* out) {
  switch (b) {
#define TQO_CASE(n) case n: bp4x_unpack_sse<n>(in, out); break;
    TQO_CASE(0) TQO_CASE(1) TQO_CASE(2) TQO_CASE(3) TQO_CASE(4) TQO_CASE(5) TQO_CASE(6) TQO_CASE(7)
    TQO_CASE(8) TQO_CASE(9) TQO_CASE(10) TQO_CASE(11) TQO_CASE(12) TQO_CASE(13) TQO_CASE(14) TQO_CASE(15)
    TQO_CASE(16) TQO_CASE(17) TQO_CASE(18) TQO_CASE(19) TQO_CASE(20) TQO_CASE(21) TQO_CASE(22) TQO_CASE(23)
    TQO_CASE(24) TQO_CASE(25) TQO_CASE(26) TQO_CASE(27) TQO_CASE(28) TQO_CASE(29) TQO_CASE(30) TQO_CASE(31)
    TQO_CASE(32)
#undef TQO_CASE
    default: std::memset(out, 0, 512);
  }
}

// BitPacker::num_bits / num_bits_sorted / num_bits_strictly_sorted
inline uint8_t bp4x_num_bits(const uint32_t* v) {
  uint32_t acc = 0;
  for (int j = 0; j < 128; ++j) acc |= v[j];
  return bit_length(acc);
}
// initial_is_none: strictly-sorted with initial == None  => previous value is u32::MAX.
inline void bp4x_deltas(const uint32_t* v, bool strict, bool initial_is_none, uint32_t initial,
                        uint32_t* deltas) {
  uint32_t prev = (strict && initial_is_none) ? 0xFFFFFFFFu : initial;
  for (int j = 0; j < 128; ++j) {
    deltas[j] = v[j] - prev - (strict ? 1u : 0u);  // wrapping
    prev = v[j];
  }
}
inline void bp4x_integrate(uint32_t* vals, bool strict, bool initial_is_none, uint32_t initial) {
  // 4-lane SSE prefix sum with carry, as a SIMD decoder does it.
  uint32_t carry = (strict && initial_is_none) ? 0xFFFFFFFFu : initial;
  __m128i prev = _mm_set1_epi32((int)carry);
  const __m128i one = _mm_set1_epi32(strict ? 1 : 0);
  __m128i* p = reinterpret_cast<__m128i*>(vals);
  for (int row = 0; row < 32; ++row) {
    __m128i d = _mm_add_epi32(_mm_loadu_si128(p + row), one);
    d = _mm_add_epi32(d, _mm_slli_si128(d, 4));
    d = _mm_add_epi32(d, _mm_slli_si128(d, 8));
    d = _mm_add_epi32(d, prev);
    _mm_storeu_si128(p + row, d);
    prev = _mm_shuffle_epi32(d, 0xFF);
  }
}

// ---- postings VInt (compression/vint.rs) ---------------------------------------------------
inline size_t vint_compress_sorted(const uint32_t* input, size_t n, uint8_t* output, uint32_t offset) {
  size_t written = 0;
  for (size_t i = 0; i < n; ++i) {
    uint32_t to_encode = input[i] - offset;
    offset = input[i];
    for (;;) {
      uint8_t next_byte = (uint8_t)(to_encode % 128u);
      to_encode /= 128u;
      if (to_encode == 0) { output[written++] = next_byte | 128u; break; }
      output[written++] = next_byte;
    }
  }
  return written;
}
inline size_t vint_compress_unsorted(const uint32_t* input, size_t n, uint8_t* output) {
  size_t written = 0;
  for (size_t i = 0; i < n; ++i) {
    uint32_t to_encode = input[i];
    for (;;) {
      uint8_t next_byte = (uint8_t)(to_encode % 128u);
      to_encode /= 128u;
      if (to_encode == 0) { output[written++] = next_byte | 128u; break; }
      output[written++] = next_byte;
    }
  }
  return written;
}
inline size_t vint_uncompress_sorted(const uint8_t* data, uint32_t* output, size_t n, uint32_t offset) {
  size_t read = 0;
  uint32_t result = offset;
  for (size_t i = 0; i < n; ++i) {
    uint32_t shift = 0;
    for (;;) {
      uint8_t cur = data[read++];
      result += (uint32_t)(cur % 128u) << shift;
      if (cur & 128u) break;
      shift += 7;
    }
    output[i] = result;
  }
  return read;
}
inline size_t vint_uncompress_unsorted(const uint8_t* data, uint32_t* output, size_t n) {
  size_t read = 0;
  for (size_t i = 0; i < n; ++i) {
    uint32_t result = 0, shift = 0;
    for (;;) {
      uint8_t cur = data[read++];
      result += (uint32_t)(cur % 128u) << shift;
      if (cur & 128u) break;
      shift += 7;
    }
    output[i] = result;
  }
  return read;
}

// common/src/vint.rs: VInt(u64) serialize / deserialize (7 bits per byte, STOP_BIT on last byte)
inline void common_vint_serialize(uint64_t val, std::vector<uint8_t>& out) {
  for (;;) {
    uint8_t b = (uint8_t)(val % 128u);
    val /= 128u;
    if (val == 0) { out.push_back(b | 128u); return; }
    out.push_back(b);
  }
}
inline bool common_vint_deserialize(const uint8_t* data, size_t len, uint64_t* val, size_t* consumed) {
  uint64_t result = 0;
  uint32_t shift = 0;
  for (size_t i = 0; i < len; ++i) {
    uint8_t b = data[i];
    result |= (uint64_t)(b % 128u) << shift;
    if (b >= 128u) { *val = result; *consumed = i + 1; return true; }
    shift += 7;
    if (shift > 63) return false;
  }
  return false;
}

// ---- BlockEncoder / BlockDecoder (compression/mod.rs:17-168) -------------------------------
struct BlockEncoder {
  uint8_t output[COMPRESSION_BLOCK_SIZE * 5];
  // returns (num_bits, written size)
  uint8_t compress_block_sorted(const uint32_t* block, uint32_t offset, size_t* written) {
    const bool none = (offset == 0);  // mod.rs:36-39
    uint32_t deltas[128];
    bp4x_deltas(block, true, none, offset, deltas);
    const uint8_t num_bits = bp4x_num_bits(deltas);
    bp4x_pack(deltas, num_bits, output);
    *written = compressed_block_size(num_bits);
    return num_bits;
  }
  uint8_t compress_block_unsorted(const uint32_t* block, bool minus_one_encoded, size_t* written) {
    uint32_t tmp[128];
    const uint32_t* src = block;
    if (minus_one_encoded) {
      for (int i = 0; i < 128; ++i) tmp[i] = block[i] - 1;
      src = tmp;
    }
    const uint8_t num_bits = bp4x_num_bits(src);
    bp4x_pack(src, num_bits, output);
    *written = compressed_block_size(num_bits);
    return num_bits;
  }
  size_t compress_vint_sorted(const uint32_t* input, size_t n, uint32_t offset) {
    return vint_compress_sorted(input, n, output, offset);
  }
  size_t compress_vint_unsorted(const uint32_t* input, size_t n) {
    return vint_compress_unsorted(input, n, output);
  }
};

struct BlockDecoder {
  alignas(16) uint32_t output[COMPRESSION_BLOCK_SIZE];
  size_t output_len = 0;
  explicit BlockDecoder(uint32_t val = 0) { for (auto& o : output) o = val; }

  size_t uncompress_block_sorted(const uint8_t* data, uint32_t offset, uint8_t num_bits, bool strict_delta) {
    output_len = COMPRESSION_BLOCK_SIZE;
    bp4x_unpack(data, num_bits, output);
    if (strict_delta) bp4x_integrate(output, true, offset == 0, offset);  // mod.rs:112-113
    else bp4x_integrate(output, false, false, offset);
    return compressed_block_size(num_bits);
  }
  size_t uncompress_block_unsorted(const uint8_t* data, uint8_t num_bits, bool minus_one_encoded) {
    output_len = COMPRESSION_BLOCK_SIZE;
    bp4x_unpack(data, num_bits, output);
    if (minus_one_encoded) for (auto& o : output) o += 1;
    return compressed_block_size(num_bits);
  }
  size_t uncompress_vint_sorted(const uint8_t* data, uint32_t offset, size_t num_els, uint32_t padding) {
    output_len = num_els;
    for (auto& o : output) o = padding;
    return vint_uncompress_sorted(data, output, num_els, offset);
  }
  size_t uncompress_vint_unsorted(const uint8_t* data, size_t num_els, uint32_t padding) {
    output_len = num_els;
    for (auto& o : output) o = padding;
    return vint_uncompress_unsorted(data, output, num_els);
  }
};

// src/postings/block_search.rs:38-76 — branchless 8-ary lower bound over a padded 128 block.
inline size_t search_block(const uint32_t* arr, uint32_t target) {
  size_t base = 0, range = COMPRESSION_BLOCK_SIZE;
  constexpr size_t K = 8;
  for (;;) {
    const size_t step = range / K;
    if (step == 0) break;
    size_t count = 0;
    for (size_t i = 1; i < K; ++i) count += (arr[base + i * step - 1] < target) ? 1 : 0;
    base += count * step;
    range = step;
  }
  size_t count = 0;
  for (size_t i = 0; i < range; ++i) count += (arr[base + i] < target) ? 1 : 0;
  return base + count;
}

}  // namespace tqo
