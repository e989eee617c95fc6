// Host engine behind the C ABI of include/tantivy_b200.h: segment registry in HBM, per-term
// block-table cache, batch planning, kernel launches, result fetch.
//
// Replaces, for TermQuery / all-MUST / all-SHOULD BooleanQuery of TermQuerys collected by
// TopDocs::order_by_score, the reference's per-segment loop
//   Searcher::search_with_executor            src/core/searcher.rs:220-237
//   SortBySimilarityScore::collect_segment_top_k   src/collector/sort_key/sort_by_score.rs:35-66
//   Weight::for_each_pruning                  src/query/weight.rs:123-132
//   TopBySortKeyCollector::merge_fruits       src/collector/sort_key_top_collector.rs:54-60
// There is NO CPU fallback: without a CUDA device every entry point fails with TQ_ERR_CUDA.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/tantivy_b200.h"
#include "bm25_host.hpp"
#include "segment_writer.hpp"
#include "tq_kernels.cuh"

using namespace tq;

static constexpr size_t kOrDynSmem = kTileDocs * sizeof(float) + kTileDocs;  // score slots + fieldnorm bytes

namespace {

thread_local std::string g_err;

struct DevBuf {
  uint8_t* p = nullptr;
  size_t cap = 0;
  cudaError_t ensure(size_t n) {
    if (n <= cap) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
    const size_t want = std::max<size_t>(n + n / 4, 1 << 20);
    cudaError_t e = cudaMalloc(&p, want);
    if (e == cudaSuccess) cap = want;
    return e;
  }
  void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};
struct PinBuf {
  uint8_t* p = nullptr;
  size_t cap = 0;
  cudaError_t ensure(size_t n) {
    if (n <= cap) return cudaSuccess;
    if (p) cudaFreeHost(p);
    p = nullptr;
    cap = 0;
    const size_t want = std::max<size_t>(n + n / 4, 1 << 16);
    cudaError_t e = cudaMallocHost(&p, want);
    if (e == cudaSuccess) cap = want;
    return e;
  }
  void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
};

struct Segment {
  uint32_t segment_ord, field, max_doc;
  int record_option;
  uint8_t* d_idx = nullptr;  // field body incl. the 8-byte header, padded
  size_t idx_len = 0;
  uint8_t* d_fieldnorm = nullptr;
  uint8_t* d_alive = nullptr;
};

struct ListKey {
  uint32_t segment_ord, field;
  uint64_t postings_start;  // bit 63: the table was built with term frequencies ignored (TQ_TERM_IGNORE_FREQ)
  bool operator==(const ListKey& o) const { return segment_ord == o.segment_ord && field == o.field && postings_start == o.postings_start; }
};
struct ListKeyHash {
  size_t operator()(const ListKey& k) const {
    uint64_t h = k.postings_start * 0x9E3779B97F4A7C15ull ^ ((uint64_t)k.segment_ord << 32 | k.field) * 0xC2B2AE3D27D4EB4Full;
    return (size_t)(h ^ (h >> 29));
  }
};

// Bump allocator over big cudaMalloc chunks for the per-term block tables (immutable, never freed
// individually; dropped with the context).
struct Arena {
  std::vector<uint8_t*> chunks;
  size_t chunk_size = 64u << 20, used = 0;
  uint8_t* alloc(size_t n, cudaError_t* err) {
    n = (n + 255) & ~(size_t)255;
    if (chunks.empty() || used + n > chunk_size) {
      const size_t sz = std::max(chunk_size, n);
      uint8_t* p = nullptr;
      *err = cudaMalloc(&p, sz);
      if (*err != cudaSuccess) return nullptr;
      chunks.push_back(p);
      used = 0;
      if (sz > chunk_size) { used = sz; return p; }
    }
    uint8_t* r = chunks.back() + used;
    used += n;
    *err = cudaSuccess;
    return r;
  }
  void release() { for (auto* c : chunks) cudaFree(c); chunks.clear(); used = 0; }
};

uint32_t env_u32(const char* name, uint32_t def) {
  const char* v = getenv(name);
  if (!v || !*v) return def;
  return (uint32_t)strtoul(v, nullptr, 10);
}

}  // namespace

struct tq_ctx {
  int device = 0;
  std::mutex mu;  // guards segments, list cache, arena, batch pool, stats
  std::map<std::pair<uint32_t, uint32_t>, Segment> segments;
  ListDesc* d_lists = nullptr;
  uint32_t lists_cap = 0, n_lists = 0;
  std::unordered_map<ListKey, uint32_t, ListKeyHash> list_cache;
  Arena arena;
  cudaStream_t build_stream = nullptr;
  PinBuf build_pin;
  DevBuf build_dev;
  std::vector<tq_batch*> pool;
  tq_stats stats{};
  uint32_t term_blocks_per_unit, and_blocks_per_unit, or_tiles_per_unit;
  unsigned long long* d_counters = nullptr;
  uint32_t or_prune = 1, or_strip = 1, or_pipe = 1, strip_prune = 1, strip_sample_div = 32, strip_sample_div2 = 8, strip_sample_div3 = 2, strip_ne_div = 8, strip_ne_div2 = 64;
};

struct tq_batch {
  tq_ctx* ctx = nullptr;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev_start = nullptr, ev_k0 = nullptr, ev_k1 = nullptr, ev_end = nullptr;
  cudaEvent_t ev_op[4] = {nullptr, nullptr, nullptr, nullptr};  // after term / and / or / final
  PinBuf pin;      // staged descriptors (H2D source)
  DevBuf dev;      // descriptors on device
  DevBuf scratch;  // qstate + candidates + results
  PinBuf res_pin;  // results (D2H target)
  BatchParams params{};
  size_t desc_bytes = 0;
  uint32_t nq = 0, kmax = 0;
  uint32_t n_units[7] = {0, 0, 0, 0, 0, 0, 0};  // term, and, or (window kernel), or (strip kernel), strip threshold rounds 1..3
  uint32_t unit_base[7] = {0, 0, 0, 0, 0, 0, 0};
  uint32_t strip_cached_max = 0;
  uint32_t or_max_lists = 0;  // most clauses of any window-kernel union in the batch
  size_t qinit_off = 0;
  int next_phase = 0;  // of the current run (0: none started)
  size_t qstate_off = 0, cands_off = 0, res_off = 0, res_bytes = 0, n_cands = 0;
  tq_stats stats{};
  bool ran = false;
};

#define TQ_CUDA(expr)                                                                         \
  do {                                                                                        \
    cudaError_t _e = (expr);                                                                  \
    if (_e != cudaSuccess) {                                                                  \
      g_err = std::string(#expr) + ": " + cudaGetErrorString(_e);                             \
      return TQ_ERR_CUDA;                                                                     \
    }                                                                                         \
  } while (0)

static int fail(int code, const std::string& msg) { g_err = msg; return code; }

// host twin of score_to_key (tq_device.cuh): order-preserving u32 image of a float
static uint32_t host_score_key(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  return u ^ ((u >> 31) ? 0xFFFFFFFFu : 0x80000000u);
}

extern "C" {

const char* tq_last_error(tq_ctx*) { return g_err.c_str(); }

int tq_ctx_create(int device, tq_ctx** out) {
  if (!out) return fail(TQ_ERR_INVALID_ARGUMENT, "out is null");
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0)
    return fail(TQ_ERR_CUDA, std::string("no CUDA device: the B200 path has no CPU fallback (") + cudaGetErrorString(e) + ")");
  if (device < 0 || device >= n) return fail(TQ_ERR_INVALID_ARGUMENT, "bad device ordinal");
  TQ_CUDA(cudaSetDevice(device));
  auto* c = new tq_ctx();
  c->device = device;
  c->lists_cap = env_u32("TQ_MAX_LISTS", 1u << 20);
  c->term_blocks_per_unit = env_u32("TQ_TERM_BLOCKS_PER_UNIT", 512);
  c->and_blocks_per_unit = env_u32("TQ_AND_BLOCKS_PER_UNIT", 128);
  c->or_tiles_per_unit = env_u32("TQ_OR_TILES_PER_UNIT", 16);
  c->or_strip = env_u32("TQ_OR_STRIP", 1);
  c->or_pipe = env_u32("TQ_OR_PIPE", 1);
  c->strip_sample_div = env_u32("TQ_STRIP_SAMPLE_DIV", 32);  // share of a pair's windows in the threshold sample (0/1: off)
  c->strip_sample_div2 = env_u32("TQ_STRIP_SAMPLE_DIV2", 8);  // second sample round ends at this share (0/1: one round only)
  c->strip_sample_div3 = env_u32("TQ_STRIP_SAMPLE_DIV3", 2);  // third round: up to half of the windows
  c->strip_ne_div = env_u32("TQ_STRIP_NE_DIV", 8);
  c->strip_ne_div2 = env_u32("TQ_STRIP_NE_DIV2", 64);
  c->strip_prune = env_u32("TQ_STRIP_PRUNE", 1);  // MaxScore split inside k_or_strip (exact)
  c->or_prune = env_u32("TQ_OR_PRUNE", 0);  // MaxScore route: exact, but only pays off for small k / rare terms
  cudaError_t err = cudaMalloc(&c->d_lists, (size_t)c->lists_cap * sizeof(ListDesc));
  if (err == cudaSuccess) err = cudaMalloc(&c->d_counters, 8 * sizeof(unsigned long long));
  if (err == cudaSuccess) err = cudaMemset(c->d_counters, 0, 8 * sizeof(unsigned long long));
  if (err == cudaSuccess) err = cudaStreamCreateWithFlags(&c->build_stream, cudaStreamNonBlocking);
  if (err == cudaSuccess) err = cudaFuncSetAttribute(k_or, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kOrDynSmem);
  if (err == cudaSuccess) err = cudaFuncSetAttribute(k_or_pipe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pipe_smem_bytes());
  if (err == cudaSuccess) err = cudaFuncSetAttribute(k_or_strip, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)strip_smem_bytes(kMaxCached));
  if (err != cudaSuccess) { delete c; return fail(TQ_ERR_CUDA, cudaGetErrorString(err)); }
  *out = c;
  return TQ_OK;
}

void tq_batch_destroy_real(tq_batch* b) {
  if (!b) return;
  cudaSetDevice(b->ctx->device);
  if (b->stream) cudaStreamSynchronize(b->stream);
  b->pin.release(); b->dev.release(); b->scratch.release(); b->res_pin.release();
  if (b->ev_start) cudaEventDestroy(b->ev_start);
  if (b->ev_k0) cudaEventDestroy(b->ev_k0);
  if (b->ev_k1) cudaEventDestroy(b->ev_k1);
  if (b->ev_end) cudaEventDestroy(b->ev_end);
  for (auto& e : b->ev_op) if (e) cudaEventDestroy(e);
  if (b->stream) cudaStreamDestroy(b->stream);
  delete b;
}

void tq_ctx_destroy(tq_ctx* c) {
  if (!c) return;
  cudaSetDevice(c->device);
  cudaDeviceSynchronize();
  for (auto* b : c->pool) tq_batch_destroy_real(b);
  for (auto& kv : c->segments) {
    cudaFree(kv.second.d_idx); cudaFree(kv.second.d_fieldnorm); cudaFree(kv.second.d_alive);
  }
  c->arena.release();
  c->build_pin.release(); c->build_dev.release();
  if (c->build_stream) cudaStreamDestroy(c->build_stream);
  cudaFree(c->d_lists);
  cudaFree(c->d_counters);
  delete c;
}

int tq_get_stats(tq_ctx* c, tq_stats* out) {
  if (!c || !out) return fail(TQ_ERR_INVALID_ARGUMENT, "null");
  std::lock_guard<std::mutex> g(c->mu);
  *out = c->stats;
  out->lists_cached = c->n_lists;
  cudaSetDevice(c->device);
  unsigned long long h[8];
  if (cudaMemcpy(h, c->d_counters, sizeof(h), cudaMemcpyDeviceToHost) == cudaSuccess)
    for (int i = 0; i < 8; ++i) out->or_windows[i] = h[i];
  return TQ_OK;
}

int tq_segment_register(tq_ctx* c, uint32_t segment_ord, uint32_t field, uint32_t max_doc, int record_option,
                        const uint8_t* idx_body, size_t idx_len, const uint8_t* fieldnorm, size_t fieldnorm_len,
                        const uint8_t* alive_bitset, size_t alive_len) {
  if (!c || !idx_body || idx_len < 8) return fail(TQ_ERR_INVALID_ARGUMENT, "idx_body must hold the 8-byte header");
  if (record_option < 0 || record_option > 2) return fail(TQ_ERR_INVALID_ARGUMENT, "record_option");
  if (max_doc >= TQ_TERMINATED) return fail(TQ_ERR_INVALID_ARGUMENT, "max_doc");
  if (fieldnorm && fieldnorm_len < max_doc) return fail(TQ_ERR_INVALID_ARGUMENT, "fieldnorm shorter than max_doc");
  if (alive_bitset && alive_len * 8 < max_doc) return fail(TQ_ERR_INVALID_ARGUMENT, "alive bitset shorter than max_doc");
  TQ_CUDA(cudaSetDevice(c->device));
  std::lock_guard<std::mutex> g(c->mu);
  if (c->segments.count({segment_ord, field})) return fail(TQ_ERR_INVALID_ARGUMENT, "segment/field already registered");
  Segment s;
  s.segment_ord = segment_ord; s.field = field; s.max_doc = max_doc; s.record_option = record_option; s.idx_len = idx_len;
  const size_t pad = 256;  // decode_block reads one word past a block; the aligned block copy of k_build_tables reads 64 + 8 bytes past the last block
  TQ_CUDA(cudaMalloc(&s.d_idx, idx_len + pad));
  TQ_CUDA(cudaMemset(s.d_idx + idx_len, 0, pad));
  TQ_CUDA(cudaMemcpy(s.d_idx, idx_body, idx_len, cudaMemcpyHostToDevice));
  if (fieldnorm) {
    const size_t padded = ((size_t)max_doc + kTileDocs - 1) / kTileDocs * kTileDocs + kTileDocs;  // k_or stages whole windows
    TQ_CUDA(cudaMalloc(&s.d_fieldnorm, padded));
    TQ_CUDA(cudaMemset(s.d_fieldnorm, 0, padded));
    TQ_CUDA(cudaMemcpy(s.d_fieldnorm, fieldnorm, max_doc, cudaMemcpyHostToDevice));
  }
  if (alive_bitset) {
    const size_t alive_padded = ((alive_len + 7) & ~(size_t)7) + 8;  // k_count reads whole 32-bit words
    TQ_CUDA(cudaMalloc(&s.d_alive, alive_padded));
    TQ_CUDA(cudaMemset(s.d_alive, 0, alive_padded));
    TQ_CUDA(cudaMemcpy(s.d_alive, alive_bitset, alive_len, cudaMemcpyHostToDevice));
  }
  c->segments[{segment_ord, field}] = s;
  return TQ_OK;
}

int tq_segment_unregister(tq_ctx* c, uint32_t segment_ord, uint32_t field) {
  if (!c) return fail(TQ_ERR_INVALID_ARGUMENT, "null");
  cudaSetDevice(c->device);
  std::lock_guard<std::mutex> g(c->mu);
  auto it = c->segments.find({segment_ord, field});
  if (it == c->segments.end()) return fail(TQ_ERR_NOT_FOUND, "segment/field not registered");
  cudaDeviceSynchronize();
  cudaFree(it->second.d_idx); cudaFree(it->second.d_fieldnorm); cudaFree(it->second.d_alive);
  c->segments.erase(it);
  for (auto li = c->list_cache.begin(); li != c->list_cache.end();)
    if (li->first.segment_ord == segment_ord && li->first.field == field) li = c->list_cache.erase(li); else ++li;
  return TQ_OK;
}

}  // extern "C"

// ---- list cache ---------------------------------------------------------------------------------
namespace {

struct PendingBuild { BuildJob job; ListDesc desc; };

// Looks a posting list up in the cache or schedules its table build. ctx->mu held.
int get_list(tq_ctx* c, const tq_term_seg& ts, bool ignore_freq, std::vector<PendingBuild>& pending, uint32_t* list_id, const Segment** seg_out) {
  auto sit = c->segments.find({ts.segment_ord, ts.field});
  if (sit == c->segments.end()) return fail(TQ_ERR_NOT_FOUND, "term_seg names a segment/field that is not registered");
  const Segment& seg = sit->second;
  *seg_out = &seg;
  if (ts.postings_end < ts.postings_start || ts.postings_end + 8 > seg.idx_len) return fail(TQ_ERR_INVALID_ARGUMENT, "postings range outside the field body");
  if (ts.postings_end - ts.postings_start > 0xFFFFFFFFull) return fail(TQ_ERR_UNSUPPORTED, "posting list larger than 4 GiB");
  ignore_freq = ignore_freq && seg.record_option != 0;
  const ListKey key{ts.segment_ord, ts.field, ts.postings_start | (ignore_freq ? 1ull << 63 : 0ull)};
  auto it = c->list_cache.find(key);
  if (it != c->list_cache.end()) { *list_id = it->second; return TQ_OK; }
  if (c->n_lists >= c->lists_cap) return fail(TQ_ERR_OOM, "posting-list table cache full (TQ_MAX_LISTS)");
  const uint32_t n_blocks = ts.doc_freq / 128u, tail_n = ts.doc_freq % 128u;
  cudaError_t e;
  const size_t n_last = (size_t)n_blocks + 1, n_blk = (size_t)n_blocks + 1;
  const size_t len = (size_t)(ts.postings_end - ts.postings_start);
  const size_t copy_bytes = ((len + 15) & ~(size_t)15) + 128;  // 16-byte aligned copy of the blocks + slack
  uint8_t* mem = c->arena.alloc(copy_bytes + n_last * 16 + n_last * 4 + 12 + n_blk * 8 + (size_t)tail_n * 8 + 16, &e);
  if (!mem) return fail(TQ_ERR_OOM, std::string("block table alloc: ") + cudaGetErrorString(e));
  PendingBuild pb;
  ListDesc& d = pb.desc;
  memset(&d, 0, sizeof(d));
  uint8_t* p = mem;
  d.blocks = p; p += copy_bytes;  // filled by k_build_tables
  d.tab4 = reinterpret_cast<const uint4*>(p); p += n_last * 16;
  d.blk = reinterpret_cast<const uint2*>(p); p += n_blk * 8;
  d.last_doc = reinterpret_cast<const uint32_t*>(p); p += n_last * 4;
  d.tail_docs = reinterpret_cast<const uint32_t*>(p); p += (size_t)tail_n * 4;
  d.tail_tfs = reinterpret_cast<const uint32_t*>(p);
  d.fieldnorm = seg.d_fieldnorm;
  d.n_blocks = n_blocks; d.tail_n = tail_n; d.n_total = n_blocks + (tail_n ? 1u : 0u); d.doc_freq = ts.doc_freq;
  pb.job.bytes = seg.d_idx + 8 + ts.postings_start;
  pb.job.len = (uint32_t)(ts.postings_end - ts.postings_start);
  pb.job.doc_freq = ts.doc_freq;
  pb.job.record_option = (uint32_t)seg.record_option | (ignore_freq ? 0x100u : 0u);
  pb.job.list_id = c->n_lists;
  *list_id = c->n_lists++;
  c->list_cache.emplace(key, *list_id);
  pending.push_back(pb);
  return TQ_OK;
}

// Builds every pending table and waits for it (first use of a term only). ctx->mu held.
int flush_builds(tq_ctx* c, std::vector<PendingBuild>& pending, uint64_t* built) {
  if (pending.empty()) return TQ_OK;
  const size_t n = pending.size();
  // new lists have consecutive ids
  const uint32_t first_id = pending.front().job.list_id;
  TQ_CUDA(c->build_pin.ensure(n * (sizeof(ListDesc) + sizeof(BuildJob))));
  TQ_CUDA(c->build_dev.ensure(n * sizeof(BuildJob)));
  ListDesc* hd = reinterpret_cast<ListDesc*>(c->build_pin.p);
  BuildJob* hj = reinterpret_cast<BuildJob*>(c->build_pin.p + n * sizeof(ListDesc));
  for (size_t i = 0; i < n; ++i) { hd[i] = pending[i].desc; hj[i] = pending[i].job; }
  TQ_CUDA(cudaMemcpyAsync(c->d_lists + first_id, hd, n * sizeof(ListDesc), cudaMemcpyHostToDevice, c->build_stream));
  TQ_CUDA(cudaMemcpyAsync(c->build_dev.p, hj, n * sizeof(BuildJob), cudaMemcpyHostToDevice, c->build_stream));
  k_build_tables<<<(unsigned)n, kThreads, 0, c->build_stream>>>(reinterpret_cast<const BuildJob*>(c->build_dev.p), c->d_lists);
  TQ_CUDA(cudaGetLastError());
  TQ_CUDA(cudaMemcpyAsync(hd, c->d_lists + first_id, n * sizeof(ListDesc), cudaMemcpyDeviceToHost, c->build_stream));
  TQ_CUDA(cudaStreamSynchronize(c->build_stream));
  for (size_t i = 0; i < n; ++i)
    if (hd[i].build_status != 0) return fail(TQ_ERR_CORRUPT, "posting list bytes are not a valid tantivy posting list");
  *built += n;
  pending.clear();
  return TQ_OK;
}

struct CacheKey {
  std::vector<float> table;
};

}  // namespace

// ---- batches ---------------------------------------------------------------------------------------
static void collect_times(tq_batch* b) {
  float ms = 0;
  if (cudaEventElapsedTime(&ms, b->ev_k0, b->ev_k1) == cudaSuccess) b->stats.kernel_ms = ms;
  if (cudaEventElapsedTime(&ms, b->ev_k0, b->ev_op[0]) == cudaSuccess) b->stats.term_ms = ms;
  if (cudaEventElapsedTime(&ms, b->ev_op[0], b->ev_op[1]) == cudaSuccess) b->stats.and_ms = ms;
  if (cudaEventElapsedTime(&ms, b->ev_op[1], b->ev_op[2]) == cudaSuccess) b->stats.or_ms = ms;
  if (cudaEventElapsedTime(&ms, b->ev_op[2], b->ev_op[3]) == cudaSuccess) b->stats.final_ms = ms;
}

static tq_batch* acquire_batch(tq_ctx* c) {
  {
    std::lock_guard<std::mutex> g(c->mu);
    if (!c->pool.empty()) { tq_batch* b = c->pool.back(); c->pool.pop_back(); return b; }
  }
  auto* b = new tq_batch();
  b->ctx = c;
  if (cudaStreamCreateWithFlags(&b->stream, cudaStreamNonBlocking) != cudaSuccess || cudaEventCreate(&b->ev_start) != cudaSuccess ||
      cudaEventCreate(&b->ev_k0) != cudaSuccess || cudaEventCreate(&b->ev_k1) != cudaSuccess || cudaEventCreate(&b->ev_end) != cudaSuccess ||
      cudaEventCreate(&b->ev_op[0]) != cudaSuccess || cudaEventCreate(&b->ev_op[1]) != cudaSuccess || cudaEventCreate(&b->ev_op[2]) != cudaSuccess ||
      cudaEventCreate(&b->ev_op[3]) != cudaSuccess) {
    tq_batch_destroy_real(b);
    return nullptr;
  }
  return b;
}

extern "C" {

void tq_batch_destroy(tq_batch* b) {
  if (!b) return;
  cudaSetDevice(b->ctx->device);
  cudaStreamSynchronize(b->stream);
  std::lock_guard<std::mutex> g(b->ctx->mu);
  b->ctx->pool.push_back(b);  // buffers are recycled by the next batch
}

int tq_batch_prepare(tq_ctx* c, const tq_query* queries, size_t nq, tq_batch** out) {
  if (!c || !out || (!queries && nq)) return fail(TQ_ERR_INVALID_ARGUMENT, "null");
  TQ_CUDA(cudaSetDevice(c->device));
  tq_batch* b = acquire_batch(c);
  if (!b) return fail(TQ_ERR_CUDA, "stream/event creation failed");
  struct Guard { tq_batch* b; bool ok = false; ~Guard() { if (!ok) tq_batch_destroy(b); } } guard{b};
  b->ran = false;
  b->nq = (uint32_t)nq;
  b->stats = tq_stats{};

  std::vector<QList> qlists;
  std::vector<QSeg> qsegs;
  std::vector<Unit> units[7];
  std::vector<DQuery> dq(nq);
  std::vector<float> caches;  // n_caches * 256
  std::unordered_map<uint32_t, uint32_t> cache_by_avg;  // avg bits -> cache idx
  std::vector<PendingBuild> pending;
  uint64_t built = 0, alg_bytes = 0, postings = 0, op_bytes[3] = {0, 0, 0};
  uint32_t kmax = 1;
  size_t n_cands = 0;
  std::vector<int> qseg_op;
  std::vector<uint32_t> qseg_total;
  uint32_t n_qsegs_op[4] = {0, 0, 0, 0};
  std::vector<char> qseg_sample;  // strip pairs that get a threshold sample pass (MaxScore can then skip their dense clauses)
  uint32_t strip_cached_max = 0, or_max_lists = 0;
  {
    std::lock_guard<std::mutex> g(c->mu);
    std::vector<const tq_term_seg*> order;
    for (size_t qi = 0; qi < nq; ++qi) {
      const tq_query& q = queries[qi];
      if (q.k == 0 || q.k > TQ_MAX_K) return fail(TQ_ERR_INVALID_ARGUMENT, "k must be in 1..TQ_MAX_K");
      if (q.n_terms == 0 || q.n_terms > TQ_MAX_TERMS) return fail(TQ_ERR_INVALID_ARGUMENT, "n_terms must be in 1..TQ_MAX_TERMS");
      if (q.op != TQ_OP_TERM && q.op != TQ_OP_AND && q.op != TQ_OP_OR) return fail(TQ_ERR_INVALID_ARGUMENT, "op");
      if (q.op == TQ_OP_TERM && q.n_terms != 1) return fail(TQ_ERR_INVALID_ARGUMENT, "TQ_OP_TERM takes one term");
      if (!q.weight || (!q.avg_fieldnorm && !q.tf_cache) || (!q.term_segs && q.n_term_segs)) return fail(TQ_ERR_INVALID_ARGUMENT, "query arrays");
      kmax = std::max(kmax, q.k);
      alg_bytes += 12ull * q.k;
      op_bytes[q.n_terms == 1 ? TQ_OP_TERM : q.op] += 12ull * q.k;
      // tf-norm tables of this query's clauses
      uint32_t cache_idx[TQ_MAX_TERMS];
      for (uint32_t t = 0; t < q.n_terms; ++t) {
        if (q.tf_cache) {
          cache_idx[t] = (uint32_t)(caches.size() / 256);
          caches.insert(caches.end(), q.tf_cache + 256 * (size_t)t, q.tf_cache + 256 * (size_t)(t + 1));
        } else {
          uint32_t bits;
          memcpy(&bits, &q.avg_fieldnorm[t], 4);
          auto it = cache_by_avg.find(bits);
          if (it == cache_by_avg.end()) {
            float tab[256];
            bm25_tf_cache(q.avg_fieldnorm[t], tab);
            it = cache_by_avg.emplace(bits, (uint32_t)(caches.size() / 256)).first;
            caches.insert(caches.end(), tab, tab + 256);
          }
          cache_idx[t] = it->second;
        }
      }
      // effective shape: an AND / OR of one clause is that clause (boolean_weight.rs:57-68, block_wand_union.rs:154-157)
      const int op = q.n_terms == 1 ? TQ_OP_TERM : q.op;
      // group the (clause, segment) lists by segment
      order.clear();
      for (uint32_t i = 0; i < q.n_term_segs; ++i) {
        if (q.term_segs[i].term_idx >= q.n_terms) return fail(TQ_ERR_INVALID_ARGUMENT, "term_idx out of range");
        if (q.term_segs[i].doc_freq) order.push_back(&q.term_segs[i]);
      }
      std::stable_sort(order.begin(), order.end(), [](const tq_term_seg* a, const tq_term_seg* b) {
        return a->segment_ord != b->segment_ord ? a->segment_ord < b->segment_ord : a->term_idx < b->term_idx;
      });
      for (size_t i = 0; i < order.size();) {
        size_t j = i;
        while (j < order.size() && order[j]->segment_ord == order[i]->segment_ord) ++j;
        // lists of this (query, segment), clause order
        const size_t n_here = j - i;
        bool dup = false;
        for (size_t a = i + 1; a < j; ++a) dup |= order[a]->term_idx == order[a - 1]->term_idx;
        if (dup) return fail(TQ_ERR_INVALID_ARGUMENT, "duplicate (term_idx, segment_ord)");
        if (op == TQ_OP_AND && n_here != q.n_terms) { i = j; continue; }  // a clause is absent: empty intersection
        QSeg qs;
        memset(&qs, 0, sizeof(qs));
        qs.query = (uint32_t)qi;
        qs.lists_base = (uint32_t)qlists.size();
        qs.segment_ord = order[i]->segment_ord;
        uint32_t lead_total = 0;
        std::vector<std::pair<uint32_t, QList>> here;  // (doc_freq, list)
        const Segment* seg = nullptr;
        const uint8_t* fn0 = nullptr;
        bool uniform_fn = true;
        for (size_t a = i; a < j; ++a) {
          uint32_t id;
          int rc = get_list(c, *order[a], q.term_flags && (q.term_flags[order[a]->term_idx] & TQ_TERM_IGNORE_FREQ), pending, &id, &seg);
          if (rc != TQ_OK) return rc;
          if (a == i) fn0 = seg->d_fieldnorm; else uniform_fn &= (seg->d_fieldnorm == fn0);
          QList ql{id, q.weight[order[a]->term_idx], cache_idx[order[a]->term_idx], 0};
          here.push_back({order[a]->doc_freq, ql});
          alg_bytes += (order[a]->postings_end - order[a]->postings_start) + order[a]->doc_freq;
          op_bytes[op] += (order[a]->postings_end - order[a]->postings_start) + order[a]->doc_freq;
          postings += order[a]->doc_freq;
        }
        qs.max_doc = seg->max_doc;
        qs.alive = seg->d_alive;
        qs.fieldnorm = uniform_fn ? fn0 : nullptr;
        if (op == TQ_OP_AND)  // leader = rarest, then ascending doc_freq; stable (block_wand_intersection.rs:27)
          std::stable_sort(here.begin(), here.end(), [](const std::pair<uint32_t, QList>& a, const std::pair<uint32_t, QList>& b) { return a.first < b.first; });
        bool prunable = false;
        if (op == TQ_OP_AND) {
          prunable = true;
          for (auto& h : here) prunable = prunable && std::isfinite(h.second.weight) && h.second.weight >= 0.0f;
        }
        if (op == TQ_OP_OR) {
          // Canonical union order = descending Bm25Weight.weight, ties in clause order (the reference's own order is
          // data dependent, block_wand_union.rs:205-208): the f32 sum is taken in this order, and the clauses with the
          // smallest score bounds form a suffix, which is what the MaxScore split of k_or_strip needs.
          std::stable_sort(here.begin(), here.end(), [](const std::pair<uint32_t, QList>& a, const std::pair<uint32_t, QList>& b) { return a.second.weight > b.second.weight; });
          prunable = true;
          for (auto& h : here) prunable = prunable && std::isfinite(h.second.weight) && h.second.weight >= 0.0f;
        }
        qs.flags = (uniform_fn ? 1u : 0u) | (prunable ? 2u : 0u);
        int unit_class = op;
        if (op == TQ_OP_OR && c->or_strip && q.k <= kStripMaxK && here.size() <= kStripMaxLists) {
          // strip kernel: clauses with less than one block per kWin-doc window keep their current block decoded in shared memory
          uint32_t n_thin = 0;
          static const uint64_t thin_mult = env_u32("TQ_STRIP_THIN_MULT", 1u);
          auto is_thin = [&](uint32_t df) { return (uint64_t)df * (kWin / 128u) < (uint64_t)qs.max_doc * thin_mult; };
          for (auto& h : here) if (is_thin(h.first)) ++n_thin;
          if (n_thin <= kMaxCached) {
            uint32_t slot = 0;
            for (auto& h : here) h.second.pad = is_thin(h.first) ? (1u | (slot++ << 1)) : 0u;
            strip_cached_max = std::max(strip_cached_max, n_thin);
            unit_class = 3;
          }
        }
        if (unit_class == TQ_OP_OR) or_max_lists = std::max<uint32_t>(or_max_lists, (uint32_t)here.size());
        for (auto& h : here) qlists.push_back(h.second);
        qs.n_lists = (uint32_t)here.size();
        lead_total = here[0].first / 128u + ((here[0].first % 128u) ? 1u : 0u);
        qsegs.push_back(qs);
        qseg_op.push_back(unit_class);
        {
          bool any_thick = false;
          for (auto& h : here) any_thick = any_thick || (uint64_t)h.first * std::max(c->strip_ne_div, c->strip_ne_div2) >= qs.max_doc;
          qseg_sample.push_back(unit_class == 3 && prunable && any_thick);
        }
        qseg_total.push_back(unit_class == 3 ? (qs.max_doc + kWin - 1) / kWin : (op == TQ_OP_OR ? (qs.max_doc + kTileDocs - 1) / kTileDocs : lead_total));
        ++n_qsegs_op[unit_class];
        i = j;
      }
      dq[qi].k = q.k;
      dq[qi].op = (uint32_t)op;
    }
    // Work units. A unit is one CTA's share of a (query, segment). With few (query, segment) pairs in the
    // batch every pair is cut into many units (latency); with many, units grow so that a CTA's local
    // top-k threshold gets tight and few candidates reach k_final (throughput).
    const uint32_t target_units = env_u32("TQ_TARGET_UNITS", 148u * 4u * 32u);
    std::vector<size_t> q_cands(nq, 0);
    for (size_t s = 0; s < qsegs.size(); ++s) {
      const int op = qseg_op[s];
      const uint32_t total = qseg_total[s];
      const uint32_t min_per = op == TQ_OP_TERM ? c->term_blocks_per_unit : (op == TQ_OP_AND ? c->and_blocks_per_unit : (op == 3 ? kStripWarps * 64u : c->or_tiles_per_unit));
      const uint32_t want_units = std::max<uint32_t>(1u, (target_units + n_qsegs_op[op] - 1) / n_qsegs_op[op]);
      const uint32_t per = std::max<uint32_t>(min_per, (total + want_units - 1) / want_units);
      const uint32_t k = dq[qsegs[s].query].k;
      // Threshold sample: the first 1/sample_div of a strip pair's windows run in a launch of their own; the exact k-th
      // best score over all sampled windows of the query (k_theta) then seeds the threshold of the main launch, whose
      // MaxScore split drops the dense clauses from the first window on.  Nothing is scored twice.
      uint32_t first = 0;
      if (op == 3 && qseg_sample[s] && c->strip_sample_div > 1 && total >= 8u * c->strip_sample_div) {
        const uint32_t cut1 = std::max<uint32_t>(kStripWarps, total / c->strip_sample_div);
        const uint32_t cut2 = c->strip_sample_div2 > 1 && c->strip_sample_div2 < c->strip_sample_div ? std::max(cut1, total / c->strip_sample_div2) : cut1;
        const uint32_t cut3 = c->strip_sample_div3 > 1 && c->strip_sample_div3 < c->strip_sample_div2 ? std::max(cut2, total / c->strip_sample_div3) : cut2;
        const uint32_t cuts[4] = {0, cut1, cut2, cut3};
        for (int r = 0; r < 3; ++r)  // round r covers [cuts[r], cuts[r+1]); a k_theta pass follows each round
          for (uint32_t b0 = cuts[r]; b0 < cuts[r + 1]; b0 += per) {
            units[4 + r].push_back(Unit{(uint32_t)s, b0, std::min(cuts[r + 1], b0 + per), 0});
            q_cands[qsegs[s].query] += (size_t)kStripWarps * k;
          }
        first = cut3;
      }
      for (uint32_t b0 = first; b0 < total; b0 += per) {
        units[op].push_back(Unit{(uint32_t)s, b0, std::min(total, b0 + per), 0});
        q_cands[qsegs[s].query] += op == 3 ? (size_t)kStripWarps * k : 2u * (size_t)k;  // what one unit may hand over
      }
    }
    for (size_t qi = 0; qi < nq; ++qi) {
      dq[qi].cand_base = (uint32_t)n_cands;
      dq[qi].cand_cap = (uint32_t)q_cands[qi];
      n_cands += q_cands[qi];
      if (n_cands > 0xFFFFFFF0ull) return fail(TQ_ERR_UNSUPPORTED, "batch too large: split it");
    }
    int rc = flush_builds(c, pending, &built);
    if (rc != TQ_OK) return rc;
  }
  if (caches.empty()) caches.resize(256, 0.0f);

  // ---- stage descriptors -------------------------------------------------------------------------
  auto align = [](size_t x) { return (x + 255) & ~(size_t)255; };
  size_t off = 0;
  const size_t o_caches = off; off = align(off + caches.size() * 4);
  const size_t n_caches = caches.size() / 256;
  const size_t o_tftab = off; off = align(off + n_caches * kTfRows * 256 * 4);  // device only (built by k_build_tf_tables)
  const size_t o_qlists = off; off = align(off + qlists.size() * sizeof(QList));
  const size_t o_qsegs = off; off = align(off + qsegs.size() * sizeof(QSeg));
  const size_t n_units_total = units[0].size() + units[1].size() + units[2].size() + units[3].size() + units[4].size() + units[5].size() + units[6].size();
  b->strip_cached_max = strip_cached_max;
  b->or_max_lists = or_max_lists;
  const size_t o_units = off; off = align(off + n_units_total * sizeof(Unit));
  const size_t o_queries = off; off = align(off + dq.size() * sizeof(DQuery));
  const size_t o_qinit = off; off = align(off + std::max<size_t>(nq, 1) * sizeof(QState));  // per-run initial state (threshold keys)
  b->desc_bytes = off;
  TQ_CUDA(b->pin.ensure(off + 256));
  TQ_CUDA(b->dev.ensure(off + 256));
  memcpy(b->pin.p + o_caches, caches.data(), caches.size() * 4);
  if (!qlists.empty()) memcpy(b->pin.p + o_qlists, qlists.data(), qlists.size() * sizeof(QList));
  if (!qsegs.empty()) memcpy(b->pin.p + o_qsegs, qsegs.data(), qsegs.size() * sizeof(QSeg));
  {
    Unit* u = reinterpret_cast<Unit*>(b->pin.p + o_units);
    uint32_t base = 0;
    for (int op = 0; op < 7; ++op) {
      b->unit_base[op] = base;
      b->n_units[op] = (uint32_t)units[op].size();
      if (!units[op].empty()) memcpy(u + base, units[op].data(), units[op].size() * sizeof(Unit));
      base += (uint32_t)units[op].size();
    }
  }
  if (!dq.empty()) memcpy(b->pin.p + o_queries, dq.data(), dq.size() * sizeof(DQuery));
  {
    QState* qi0 = reinterpret_cast<QState*>(b->pin.p + o_qinit);
    for (size_t qi = 0; qi < std::max<size_t>(nq, 1); ++qi) {
      qi0[qi].theta = 0; qi0[qi].cand_count = 0;
      if (qi < nq && (queries[qi].flags & TQ_QUERY_HAS_THRESHOLD)) {
        // collect score > threshold: the smallest accepted key is the next representable score (NaN: no filter)
        float th = queries[qi].threshold;
        if (th == 0.0f) th = 0.0f;  // -0.0 and +0.0 are the same threshold
        if (th == th) { const uint32_t key = host_score_key(th); qi0[qi].theta = key == 0xFFFFFFFFu ? key : key + 1u; }
      }
    }
  }
  b->qinit_off = o_qinit;

  // ---- scratch: qstate | candidates | results ------------------------------------------------------
  b->kmax = kmax;
  b->n_cands = n_cands;
  size_t so = 0;
  b->qstate_off = so; so = align(so + std::max<size_t>(nq, 1) * sizeof(QState));
  b->cands_off = so; so = align(so + std::max<size_t>(n_cands, 1) * sizeof(Cand));
  b->res_off = so;
  const size_t rows = std::max<size_t>(nq, 1) * kmax;
  const size_t o_rs = 0, o_rg = align(rows * 4), o_rd = o_rg + align(rows * 4), o_rc = o_rd + align(rows * 4);
  b->res_bytes = o_rc + align(std::max<size_t>(nq, 1) * 4);
  so += b->res_bytes;
  TQ_CUDA(b->scratch.ensure(so));
  TQ_CUDA(b->res_pin.ensure(b->res_bytes));

  BatchParams& P = b->params;
  P.lists = c->d_lists;
  P.caches = reinterpret_cast<const float*>(b->dev.p + o_caches);
  P.tf_tables = reinterpret_cast<const float*>(b->dev.p + o_tftab);
  P.qlists = reinterpret_cast<const QList*>(b->dev.p + o_qlists);
  P.qsegs = reinterpret_cast<const QSeg*>(b->dev.p + o_qsegs);
  P.units = reinterpret_cast<const Unit*>(b->dev.p + o_units);
  P.queries = reinterpret_cast<const DQuery*>(b->dev.p + o_queries);
  P.qstate = reinterpret_cast<QState*>(b->scratch.p + b->qstate_off);
  P.cands = reinterpret_cast<Cand*>(b->scratch.p + b->cands_off);
  uint8_t* r = b->scratch.p + b->res_off;
  P.res_scores = reinterpret_cast<float*>(r + o_rs);
  P.res_segs = reinterpret_cast<uint32_t*>(r + o_rg);
  P.res_docs = reinterpret_cast<uint32_t*>(r + o_rd);
  P.res_counts = reinterpret_cast<uint32_t*>(r + o_rc);
  P.res_stride = kmax;
  P.n_queries = (uint32_t)nq;
  P.counters = c->d_counters;
  P.or_prune = c->or_prune;
  P.strip_prune = c->strip_prune;
  P.strip_ne_div = c->strip_ne_div;
  P.strip_ne_div2 = c->strip_ne_div2;

  TQ_CUDA(cudaEventRecord(b->ev_start, b->stream));
  TQ_CUDA(cudaMemcpyAsync(b->dev.p, b->pin.p, b->desc_bytes, cudaMemcpyHostToDevice, b->stream));
  {
    const unsigned n = (unsigned)(n_caches * kTfRows * 256);
    k_build_tf_tables<<<(n + 255) / 256, 256, 0, b->stream>>>(P.caches, reinterpret_cast<float*>(b->dev.p + o_tftab), (uint32_t)n_caches);
    TQ_CUDA(cudaGetLastError());
  }
  b->stats.lists_built = built;
  b->stats.units = n_units_total;
  b->stats.h2d_bytes = b->desc_bytes;
  b->stats.algorithmic_bytes = alg_bytes;
  b->stats.postings = postings;
  b->stats.units_term = units[0].size(); b->stats.units_and = units[1].size(); b->stats.units_or = units[2].size() + units[3].size() + units[4].size() + units[5].size() + units[6].size(); b->stats.units_or_strip = units[3].size() + units[4].size() + units[5].size() + units[6].size();
  b->stats.bytes_term = op_bytes[0]; b->stats.bytes_and = op_bytes[1]; b->stats.bytes_or = op_bytes[2];
  guard.ok = true;
  *out = b;
  return TQ_OK;
}

// Phases of a run: 0 = term / AND / window-union kernels + the unions' first threshold round, 1 and 2 = the second
// and third threshold rounds, 3 = the unions' main launch + k_final.  Sharded callers exchange thresholds in between.
constexpr int kPhases = 4;

static int run_phase(tq_batch* b, int phase) {
  TQ_CUDA(cudaSetDevice(b->ctx->device));
  const BatchParams& P = b->params;
  uint64_t launches = 0;
  if (phase != b->next_phase) return fail(TQ_ERR_INVALID_ARGUMENT, "phases run in order, each once per run");
  if (phase == 0) {
    TQ_CUDA(cudaMemcpyAsync(P.qstate, b->dev.p + b->qinit_off, std::max<size_t>(b->nq, 1) * sizeof(QState), cudaMemcpyDeviceToDevice, b->stream));
    TQ_CUDA(cudaEventRecord(b->ev_k0, b->stream));
    if (b->n_units[TQ_OP_TERM]) { k_term<<<b->n_units[TQ_OP_TERM], kThreads, 0, b->stream>>>(P, b->unit_base[TQ_OP_TERM]); ++launches; }
    TQ_CUDA(cudaEventRecord(b->ev_op[0], b->stream));
    if (b->n_units[TQ_OP_AND]) { k_and<<<b->n_units[TQ_OP_AND], kThreads, 0, b->stream>>>(P, b->unit_base[TQ_OP_AND]); ++launches; }
    TQ_CUDA(cudaEventRecord(b->ev_op[1], b->stream));
    if (b->n_units[TQ_OP_OR]) {
      // window unions: the TMA/mbarrier pipeline when every union has few enough clauses for its tables, else the plain kernel
      if (b->ctx->or_pipe && b->or_max_lists <= kPipeMaxLists && !b->ctx->or_prune)
        k_or_pipe<<<b->n_units[TQ_OP_OR], kPipeThreads, pipe_smem_bytes(), b->stream>>>(P, b->unit_base[TQ_OP_OR]);
      else
        k_or<<<b->n_units[TQ_OP_OR], kThreads, kOrDynSmem, b->stream>>>(P, b->unit_base[TQ_OP_OR]);
      ++launches;
    }
    b->stats.kernel_launches = 0;
  }
  if (phase < kPhases - 1) {  // threshold round `phase`: its windows, then the exact k-th best so far per query
    const int r = 4 + phase;
    if (b->n_units[r]) {
      k_or_strip<<<b->n_units[r], kStripThreads, strip_smem_bytes(b->strip_cached_max), b->stream>>>(P, b->unit_base[r], b->strip_cached_max);
      k_theta<<<(unsigned)b->nq, kThreads, 0, b->stream>>>(P);
      launches += 2;
    }
    TQ_CUDA(cudaGetLastError());
    b->stats.kernel_launches += launches;
    b->next_phase = phase + 1;
    return TQ_OK;
  }
  if (b->n_units[3]) { k_or_strip<<<b->n_units[3], kStripThreads, strip_smem_bytes(b->strip_cached_max), b->stream>>>(P, b->unit_base[3], b->strip_cached_max); ++launches; }
  TQ_CUDA(cudaGetLastError());
  TQ_CUDA(cudaEventRecord(b->ev_op[2], b->stream));
  if (b->nq) { k_final<<<b->nq, kThreads, 0, b->stream>>>(P); ++launches; }
  TQ_CUDA(cudaGetLastError());
  TQ_CUDA(cudaEventRecord(b->ev_op[3], b->stream));
  TQ_CUDA(cudaEventRecord(b->ev_k1, b->stream));
  b->stats.kernel_launches += launches;
  b->next_phase = 0;
  b->ran = true;
  return TQ_OK;
}

int tq_batch_run(tq_batch* b) {
  if (!b) return fail(TQ_ERR_INVALID_ARGUMENT, "null");
  for (int p = b->next_phase; p < kPhases; ++p) {
    const int rc = run_phase(b, p);
    if (rc != TQ_OK) return rc;
  }
  return TQ_OK;
}

int tq_batch_phases(tq_batch* b) { return b ? kPhases : 0; }

int tq_batch_run_phase(tq_batch* b, int phase) {
  if (!b || phase < 0 || phase >= kPhases) return fail(TQ_ERR_INVALID_ARGUMENT, "batch / phase");
  return run_phase(b, phase);
}

int tq_batch_thresholds_export_dev(tq_batch* b, int64_t* keys_dev) {
  if (!b || !keys_dev || b->next_phase == 0) return fail(TQ_ERR_INVALID_ARGUMENT, "export needs a batch between two phases of a run");
  TQ_CUDA(cudaSetDevice(b->ctx->device));
  if (b->nq) k_theta_export<<<(unsigned)((b->nq + 255) / 256), 256, 0, b->stream>>>(b->params.qstate, reinterpret_cast<long long*>(keys_dev), (uint32_t)b->nq);
  TQ_CUDA(cudaGetLastError());
  TQ_CUDA(cudaStreamSynchronize(b->stream));  // the caller's collective runs on its own stream
  return TQ_OK;
}

int tq_batch_thresholds_import_dev(tq_batch* b, const int64_t* keys_dev) {
  if (!b || !keys_dev || b->next_phase == 0) return fail(TQ_ERR_INVALID_ARGUMENT, "import needs a batch between two phases of a run");
  TQ_CUDA(cudaSetDevice(b->ctx->device));
  if (b->nq) k_theta_import<<<(unsigned)((b->nq + 255) / 256), 256, 0, b->stream>>>(b->params.qstate, reinterpret_cast<const long long*>(keys_dev), (uint32_t)b->nq);
  TQ_CUDA(cudaGetLastError());
  return TQ_OK;
}

int tq_batch_results_dev(tq_batch* b, const float** scores_dev, const uint32_t** segment_ord_dev, const uint32_t** doc_dev,
                         const uint32_t** count_dev, uint32_t* stride) {
  if (!b || !b->ran) return fail(TQ_ERR_INVALID_ARGUMENT, "batch has not run");
  TQ_CUDA(cudaSetDevice(b->ctx->device));
  TQ_CUDA(cudaStreamSynchronize(b->stream));
  if (scores_dev) *scores_dev = b->params.res_scores;
  if (segment_ord_dev) *segment_ord_dev = b->params.res_segs;
  if (doc_dev) *doc_dev = b->params.res_docs;
  if (count_dev) *count_dev = b->params.res_counts;
  if (stride) *stride = b->kmax;
  collect_times(b);
  std::lock_guard<std::mutex> g(b->ctx->mu);
  b->ctx->stats = b->stats;
  return TQ_OK;
}

int tq_batch_results_copy_dev(tq_batch* b, float* scores_dev, uint32_t* segment_ord_dev, uint32_t* doc_dev, uint32_t* count_dev) {
  if (!b || !b->ran) return fail(TQ_ERR_INVALID_ARGUMENT, "batch has not run");
  if (!scores_dev || !segment_ord_dev || !doc_dev || !count_dev) return fail(TQ_ERR_INVALID_ARGUMENT, "null output");
  TQ_CUDA(cudaSetDevice(b->ctx->device));
  const size_t rows = (size_t)b->nq * b->kmax * 4;
  TQ_CUDA(cudaMemcpyAsync(scores_dev, b->params.res_scores, rows, cudaMemcpyDeviceToDevice, b->stream));
  TQ_CUDA(cudaMemcpyAsync(segment_ord_dev, b->params.res_segs, rows, cudaMemcpyDeviceToDevice, b->stream));
  TQ_CUDA(cudaMemcpyAsync(doc_dev, b->params.res_docs, rows, cudaMemcpyDeviceToDevice, b->stream));
  TQ_CUDA(cudaMemcpyAsync(count_dev, b->params.res_counts, (size_t)b->nq * 4, cudaMemcpyDeviceToDevice, b->stream));
  TQ_CUDA(cudaStreamSynchronize(b->stream));
  collect_times(b);
  std::lock_guard<std::mutex> g(b->ctx->mu);
  b->ctx->stats = b->stats;
  return TQ_OK;
}

int tq_batch_fetch(tq_batch* b, uint32_t out_stride, float* out_scores, uint32_t* out_segment_ord, uint32_t* out_doc, uint32_t* out_count) {
  if (!b || !b->ran) return fail(TQ_ERR_INVALID_ARGUMENT, "batch has not run");
  if (!out_scores || !out_segment_ord || !out_doc || !out_count) return fail(TQ_ERR_INVALID_ARGUMENT, "null output");
  TQ_CUDA(cudaSetDevice(b->ctx->device));
  TQ_CUDA(cudaMemcpyAsync(b->res_pin.p, b->scratch.p + b->res_off, b->res_bytes, cudaMemcpyDeviceToHost, b->stream));
  TQ_CUDA(cudaEventRecord(b->ev_end, b->stream));
  TQ_CUDA(cudaStreamSynchronize(b->stream));
  const size_t rows = std::max<size_t>(b->nq, 1) * b->kmax;
  auto align = [](size_t x) { return (x + 255) & ~(size_t)255; };
  const float* rs = reinterpret_cast<const float*>(b->res_pin.p);
  const uint32_t* rg = reinterpret_cast<const uint32_t*>(b->res_pin.p + align(rows * 4));
  const uint32_t* rd = reinterpret_cast<const uint32_t*>(b->res_pin.p + 2 * align(rows * 4));
  const uint32_t* rc = reinterpret_cast<const uint32_t*>(b->res_pin.p + 3 * align(rows * 4));
  for (uint32_t q = 0; q < b->nq; ++q) {
    const uint32_t n = std::min(std::min(rc[q], b->kmax), out_stride);
    out_count[q] = rc[q];
    memcpy(out_scores + (size_t)q * out_stride, rs + (size_t)q * b->kmax, n * 4);
    memcpy(out_segment_ord + (size_t)q * out_stride, rg + (size_t)q * b->kmax, n * 4);
    memcpy(out_doc + (size_t)q * out_stride, rd + (size_t)q * b->kmax, n * 4);
  }
  collect_times(b);
  float ms = 0;
  if (cudaEventElapsedTime(&ms, b->ev_start, b->ev_end) == cudaSuccess) b->stats.total_ms = ms;
  b->stats.d2h_bytes = b->res_bytes;
  std::lock_guard<std::mutex> g(b->ctx->mu);
  b->ctx->stats = b->stats;
  return TQ_OK;
}

int tq_search_batch(tq_ctx* c, const tq_query* queries, size_t nq, uint32_t out_stride, float* out_scores, uint32_t* out_segment_ord,
                    uint32_t* out_doc, uint32_t* out_count) {
  tq_batch* b = nullptr;
  int rc = tq_batch_prepare(c, queries, nq, &b);
  if (rc != TQ_OK) return rc;
  rc = tq_batch_run(b);
  if (rc == TQ_OK) rc = tq_batch_fetch(b, out_stride, out_scores, out_segment_ord, out_doc, out_count);
  tq_batch_destroy(b);
  return rc;
}

int tq_merge_topk_dev(tq_ctx* c, uint32_t n_lists, uint32_t nq, uint32_t stride, uint32_t k, const float* scores_dev,
                      const uint32_t* segment_ord_dev, const uint32_t* doc_dev, const uint32_t* count_dev, float* out_scores_dev,
                      uint32_t* out_segment_ord_dev, uint32_t* out_doc_dev, uint32_t* out_count_dev) {
  if (!c) return fail(TQ_ERR_INVALID_ARGUMENT, "null");
  if (k == 0 || k > TQ_MAX_K || stride < 1) return fail(TQ_ERR_INVALID_ARGUMENT, "k / stride");
  TQ_CUDA(cudaSetDevice(c->device));
  if (nq == 0) return TQ_OK;
  k_merge<<<nq, kThreads, 0, 0>>>(n_lists, nq, stride, std::min(k, stride), scores_dev, segment_ord_dev, doc_dev, count_dev, out_scores_dev,
                                  out_segment_ord_dev, out_doc_dev, out_count_dev);
  TQ_CUDA(cudaGetLastError());
  TQ_CUDA(cudaStreamSynchronize(0));
  return TQ_OK;
}

// ---- Count collector -------------------------------------------------------------------------------------------
int tq_count_batch(tq_ctx* c, const tq_query* queries, size_t nq, uint64_t* out_counts) {
  if (!c || (!queries && nq) || (!out_counts && nq)) return fail(TQ_ERR_INVALID_ARGUMENT, "null");
  TQ_CUDA(cudaSetDevice(c->device));
  for (size_t qi = 0; qi < nq; ++qi) out_counts[qi] = 0;
  if (!nq) return TQ_OK;
  tq_batch* b = acquire_batch(c);
  if (!b) return fail(TQ_ERR_CUDA, "stream/event creation failed");
  struct Guard { tq_batch* b; ~Guard() { tq_batch_destroy(b); } } guard{b};
  std::vector<uint32_t> list_ids;
  std::vector<CountSeg> segs;
  std::vector<Unit> units;
  {
    std::lock_guard<std::mutex> g(c->mu);
    std::vector<PendingBuild> pending;
    uint64_t built = 0;
    std::vector<const tq_term_seg*> order;
    for (size_t qi = 0; qi < nq; ++qi) {
      const tq_query& q = queries[qi];
      if (q.n_terms == 0 || q.n_terms > TQ_MAX_TERMS) return fail(TQ_ERR_INVALID_ARGUMENT, "n_terms must be in 1..TQ_MAX_TERMS");
      if (q.op != TQ_OP_TERM && q.op != TQ_OP_AND && q.op != TQ_OP_OR) return fail(TQ_ERR_INVALID_ARGUMENT, "op");
      if (q.op == TQ_OP_TERM && q.n_terms != 1) return fail(TQ_ERR_INVALID_ARGUMENT, "TQ_OP_TERM takes one term");
      if (!q.term_segs && q.n_term_segs) return fail(TQ_ERR_INVALID_ARGUMENT, "query arrays");
      order.clear();
      for (uint32_t i = 0; i < q.n_term_segs; ++i) {
        if (q.term_segs[i].term_idx >= q.n_terms) return fail(TQ_ERR_INVALID_ARGUMENT, "term_idx out of range");
        if (q.term_segs[i].doc_freq) order.push_back(&q.term_segs[i]);
      }
      std::stable_sort(order.begin(), order.end(), [](const tq_term_seg* a, const tq_term_seg* b) {
        return a->segment_ord != b->segment_ord ? a->segment_ord < b->segment_ord : a->term_idx < b->term_idx;
      });
      for (size_t i = 0; i < order.size();) {
        size_t j = i;
        while (j < order.size() && order[j]->segment_ord == order[i]->segment_ord) ++j;
        const uint32_t n_here = (uint32_t)(j - i);
        if (q.op == TQ_OP_AND && n_here < q.n_terms) { i = j; continue; }  // a clause without postings: empty intersection
        CountSeg cs{};
        cs.query = (uint32_t)qi; cs.lists_base = (uint32_t)list_ids.size(); cs.n_lists = n_here; cs.op = (uint32_t)q.op;
        const Segment* seg = nullptr;
        uint64_t df_single = 0;
        for (size_t a = i; a < j; ++a) {
          uint32_t id;
          int rc = get_list(c, *order[a], false, pending, &id, &seg);
          if (rc != TQ_OK) return rc;
          list_ids.push_back(id);
          df_single = order[a]->doc_freq;
        }
        cs.max_doc = seg->max_doc;
        cs.alive = seg->d_alive;
        if (n_here == 1 && !seg->d_alive) {  // TermWeight::count without deletes: the term's doc_freq (term_weight.rs:179-190)
          out_counts[qi] += df_single;
          list_ids.resize(cs.lists_base);
        } else {
          const uint32_t tiles = (cs.max_doc + kTileDocs - 1) / kTileDocs, per = 8;
          for (uint32_t t0 = 0; t0 < tiles; t0 += per) units.push_back(Unit{(uint32_t)segs.size(), t0, std::min(tiles, t0 + per), 0});
          segs.push_back(cs);
        }
        i = j;
      }
    }
    int rc = flush_builds(c, pending, &built);
    if (rc != TQ_OK) return rc;
  }
  if (units.empty()) return TQ_OK;
  auto align = [](size_t x) { return (x + 255) & ~(size_t)255; };
  size_t off = 0;
  const size_t o_ids = off; off = align(off + list_ids.size() * 4);
  const size_t o_segs = off; off = align(off + segs.size() * sizeof(CountSeg));
  const size_t o_units = off; off = align(off + units.size() * sizeof(Unit));
  const size_t o_counts = off; off = align(off + nq * 8);
  TQ_CUDA(b->pin.ensure(off + 256));
  TQ_CUDA(b->dev.ensure(off + 256));
  memcpy(b->pin.p + o_ids, list_ids.data(), list_ids.size() * 4);
  memcpy(b->pin.p + o_segs, segs.data(), segs.size() * sizeof(CountSeg));
  memcpy(b->pin.p + o_units, units.data(), units.size() * sizeof(Unit));
  memset(b->pin.p + o_counts, 0, nq * 8);
  TQ_CUDA(cudaMemcpyAsync(b->dev.p, b->pin.p, off, cudaMemcpyHostToDevice, b->stream));
  CountParams P;
  P.lists = c->d_lists;
  P.list_ids = reinterpret_cast<const uint32_t*>(b->dev.p + o_ids);
  P.segs = reinterpret_cast<const CountSeg*>(b->dev.p + o_segs);
  P.units = reinterpret_cast<const Unit*>(b->dev.p + o_units);
  P.counts = reinterpret_cast<unsigned long long*>(b->dev.p + o_counts);
  k_count<<<(unsigned)units.size(), kThreads, 0, b->stream>>>(P);
  TQ_CUDA(cudaGetLastError());
  TQ_CUDA(cudaMemcpyAsync(b->pin.p + o_counts, b->dev.p + o_counts, nq * 8, cudaMemcpyDeviceToHost, b->stream));
  TQ_CUDA(cudaStreamSynchronize(b->stream));
  const unsigned long long* dc = reinterpret_cast<const unsigned long long*>(b->pin.p + o_counts);
  for (size_t qi = 0; qi < nq; ++qi) out_counts[qi] += dc[qi];
  return TQ_OK;
}

// ---- codec-level access ---------------------------------------------------------------------------
static int resolve_single(tq_ctx* c, const tq_term_seg* list, uint32_t* id) {
  std::lock_guard<std::mutex> g(c->mu);
  std::vector<PendingBuild> pending;
  const Segment* seg;
  uint64_t built = 0;
  int rc = get_list(c, *list, false, pending, id, &seg);
  if (rc != TQ_OK) return rc;
  return flush_builds(c, pending, &built);
}

int tq_decode_postings(tq_ctx* c, const tq_term_seg* list, uint32_t* out_docs, uint32_t* out_tfs) {
  if (!c || !list || !out_docs) return fail(TQ_ERR_INVALID_ARGUMENT, "null");
  TQ_CUDA(cudaSetDevice(c->device));
  if (list->doc_freq == 0) return TQ_OK;
  uint32_t id;
  int rc = resolve_single(c, list, &id);
  if (rc != TQ_OK) return rc;
  uint32_t *d_docs = nullptr, *d_tfs = nullptr;
  const size_t bytes = (size_t)list->doc_freq * 4;
  TQ_CUDA(cudaMalloc(&d_docs, bytes));
  TQ_CUDA(cudaMalloc(&d_tfs, bytes));
  const uint32_t n_total = list->doc_freq / 128u + ((list->doc_freq % 128u) ? 1u : 0u);
  k_decode_list<<<(n_total + kWarps - 1) / kWarps, kThreads>>>(c->d_lists, id, d_docs, d_tfs);
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess) e = cudaMemcpy(out_docs, d_docs, bytes, cudaMemcpyDeviceToHost);
  if (e == cudaSuccess && out_tfs) e = cudaMemcpy(out_tfs, d_tfs, bytes, cudaMemcpyDeviceToHost);
  cudaFree(d_docs); cudaFree(d_tfs);
  if (e != cudaSuccess) return fail(TQ_ERR_CUDA, cudaGetErrorString(e));
  return TQ_OK;
}

int tq_block_table(tq_ctx* c, const tq_term_seg* list, float weight, float avg_fieldnorm, uint32_t* out_last_doc, float* out_block_max) {
  if (!c || !list || !out_last_doc || !out_block_max) return fail(TQ_ERR_INVALID_ARGUMENT, "null");
  TQ_CUDA(cudaSetDevice(c->device));
  const uint32_t n = list->doc_freq / 128u;
  if (n == 0) return TQ_OK;
  uint32_t id;
  int rc = resolve_single(c, list, &id);
  if (rc != TQ_OK) return rc;
  float tab[256];
  bm25_tf_cache(avg_fieldnorm, tab);
  float *d_cache = nullptr, *d_bm = nullptr;
  uint32_t* d_last = nullptr;
  TQ_CUDA(cudaMalloc(&d_cache, 1024));
  TQ_CUDA(cudaMalloc(&d_bm, (size_t)n * 4));
  TQ_CUDA(cudaMalloc(&d_last, (size_t)n * 4));
  cudaError_t e = cudaMemcpy(d_cache, tab, 1024, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) {
    k_block_max<<<(n + 255) / 256, 256>>>(c->d_lists, id, weight, d_cache, d_last, d_bm);
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) e = cudaMemcpy(out_last_doc, d_last, (size_t)n * 4, cudaMemcpyDeviceToHost);
  if (e == cudaSuccess) e = cudaMemcpy(out_block_max, d_bm, (size_t)n * 4, cudaMemcpyDeviceToHost);
  cudaFree(d_cache); cudaFree(d_bm); cudaFree(d_last);
  if (e != cudaSuccess) return fail(TQ_ERR_CUDA, cudaGetErrorString(e));
  return TQ_OK;
}

// ---- BM25 scalars --------------------------------------------------------------------------------------
float tq_bm25_idf(uint64_t doc_freq, uint64_t doc_count) { return bm25_idf(doc_freq, doc_count); }
float tq_bm25_weight(uint64_t doc_freq, uint64_t doc_count, float boost) { return bm25_weight(doc_freq, doc_count, boost); }
void tq_bm25_tf_cache(float avg_fieldnorm, float out[256]) { bm25_tf_cache(avg_fieldnorm, out); }
uint32_t tq_id_to_fieldnorm(uint8_t id) { return id_to_fieldnorm(id); }
uint8_t tq_fieldnorm_to_id(uint32_t fieldnorm) { return fieldnorm_to_id(fieldnorm); }

// ---- segment writer ----------------------------------------------------------------------------------
struct tq_field_writer {
  std::vector<uint8_t> fieldnorm_ids;
  FieldPostingsWriter* w = nullptr;
};

int tq_field_writer_create(int record_option, uint64_t total_num_tokens, const uint8_t* fieldnorm_ids, uint32_t max_doc, tq_field_writer** out) {
  if (!out || record_option < 0 || record_option > 2) return fail(TQ_ERR_INVALID_ARGUMENT, "args");
  auto* fw = new tq_field_writer();
  if (fieldnorm_ids) fw->fieldnorm_ids.assign(fieldnorm_ids, fieldnorm_ids + max_doc);
  fw->w = new FieldPostingsWriter(record_option, total_num_tokens, fieldnorm_ids ? fw->fieldnorm_ids.data() : nullptr, max_doc);
  *out = fw;
  return TQ_OK;
}
int tq_field_writer_add_term(tq_field_writer* fw, const uint32_t* docs, const uint32_t* tfs, uint32_t doc_freq, uint64_t* postings_start,
                             uint64_t* postings_end) {
  if (!fw || (!docs && doc_freq)) return fail(TQ_ERR_INVALID_ARGUMENT, "args");
  for (uint32_t i = 0; i < doc_freq; ++i) {
    if (i && docs[i] <= docs[i - 1]) return fail(TQ_ERR_INVALID_ARGUMENT, "docs must be strictly ascending");
    if (docs[i] >= TQ_TERMINATED) return fail(TQ_ERR_INVALID_ARGUMENT, "doc id out of range");
    if (tfs && tfs[i] == 0) return fail(TQ_ERR_INVALID_ARGUMENT, "term frequency 0");
  }
  const TermInfoOut ti = fw->w->add_term(docs, tfs, doc_freq);
  if (postings_start) *postings_start = ti.postings_start;
  if (postings_end) *postings_end = ti.postings_end;
  return TQ_OK;
}
int tq_field_writer_body(tq_field_writer* fw, const uint8_t** body, size_t* len) {
  if (!fw || !body || !len) return fail(TQ_ERR_INVALID_ARGUMENT, "args");
  *body = fw->w->body().data();
  *len = fw->w->body().size();
  return TQ_OK;
}
void tq_field_writer_destroy(tq_field_writer* fw) {
  if (!fw) return;
  delete fw->w;
  delete fw;
}

}  // extern "C"
