// Shared device/host helpers for the mpgan_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>

namespace mpg {

// ---- error plumbing (C-ABI returns int status; message kept thread-local) -------------------
void set_error(const char* fmt, ...);
#define MPG_CHECK(cond, ...)                      \
  do {                                            \
    if (!(cond)) {                                \
      ::mpg::set_error(__VA_ARGS__);              \
      return 1;                                   \
    }                                             \
  } while (0)
#define MPG_CUDA(expr)                                                              \
  do {                                                                              \
    cudaError_t _e = (expr);                                                        \
    if (_e != cudaSuccess) {                                                        \
      ::mpg::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e),      \
                       __FILE__, __LINE__);                                         \
      return 2;                                                                     \
    }                                                                               \
  } while (0)
// every kernel launch goes through here: the counter backs bench.py's "gpu_launches" claim
extern unsigned long long g_launch_count;
#define MPG_LAUNCH_CHECK()          \
  do {                              \
    ++::mpg::g_launch_count;        \
    MPG_CUDA(cudaGetLastError());   \
  } while (0)

// ---- Philox4x32-10 counter-based RNG -----------------------------------------------------------
// Dropout masks must be regenerated bit-identically in forward, recompute and backward, so they
// are a pure function of (seed, stream, row, column):  see drop_keep().
struct u4 { uint32_t x, y, z, w; };

__host__ __device__ __forceinline__ u4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                     uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint64_t p0 = (uint64_t)0xD2511F53u * c0;
    const uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
    const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
    const uint32_t n1 = (uint32_t)p1;
    const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
    const uint32_t n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  return u4{c0, c1, c2, c3};
}

// 128 keep-bits for columns [128*blk, 128*blk+128) of `row` in RNG stream `stream` (p == 0.5 path)
__host__ __device__ __forceinline__ u4 drop_bits128(uint64_t seed, uint32_t stream, uint64_t row, uint32_t blk) {
  return philox4x32_10((uint32_t)row, (uint32_t)(row >> 32), stream, blk, (uint32_t)seed, (uint32_t)(seed >> 32));
}

// Bernoulli keep decision for element (row, col).  p == 0.5: one Philox bit per element;
// otherwise a 16-bit uniform per element compared with round(p * 65536).
struct DropCfg {
  float p;          // drop probability (0 => disabled)
  float scale;      // 1/(1-p)
  uint32_t thr16;   // round(p*65536)
  int half;         // p == 0.5 fast path
  uint64_t seed;
  const uint64_t* seed_ptr;   // optional device-resident increment (CUDA-graph replay draws fresh masks)
};
__host__ __device__ __forceinline__ DropCfg make_drop(float p, uint64_t seed, const uint64_t* seed_ptr = nullptr) {
  DropCfg d;
  d.p = p; d.seed = seed; d.seed_ptr = seed_ptr;
  d.scale = p > 0.f ? 1.f / (1.f - p) : 1.f;
  d.thr16 = (uint32_t)(p * 65536.f + 0.5f);
  d.half = (p == 0.5f);
  return d;
}
// fold the device-resident seed increment in (call once at kernel entry)
__device__ __forceinline__ void resolve_seed(DropCfg& d) {
  if (d.p > 0.f && d.seed_ptr != nullptr) d.seed += *d.seed_ptr;
}
__host__ __device__ __forceinline__ bool drop_keep(const DropCfg& d, uint32_t stream, uint64_t row, uint32_t col) {
  if (d.half) {
    u4 r = drop_bits128(d.seed, stream, row, col >> 7);
    const uint32_t w = (col >> 5) & 3;
    const uint32_t word = w == 0 ? r.x : (w == 1 ? r.y : (w == 2 ? r.z : r.w));
    return (word >> (col & 31)) & 1u;
  }
  u4 r = philox4x32_10((uint32_t)row, (uint32_t)(row >> 32), stream | 0x80000000u, col >> 3,
                       (uint32_t)d.seed, (uint32_t)(d.seed >> 32));
  const uint32_t w = (col >> 1) & 3;
  const uint32_t word = w == 0 ? r.x : (w == 1 ? r.y : (w == 2 ? r.z : r.w));
  const uint32_t h = (col & 1) ? (word >> 16) : (word & 0xFFFFu);
  return h >= d.thr16;
}

// keep bits (bit j = column 32*col32 + j) of one row for the p == 0.5 path: ONE Philox draw serves 32 columns
// (same bits as drop_keep(), which evaluates them one element at a time)
__host__ __device__ __forceinline__ uint32_t drop_word32(const DropCfg& d, uint32_t stream, uint64_t row, uint32_t col32) {
  const u4 r = drop_bits128(d.seed, stream, row, col32 >> 2);
  const uint32_t w = col32 & 3;
  return w == 0 ? r.x : (w == 1 ? r.y : (w == 2 ? r.z : r.w));
}

// ---- edge-network dropout, p == 0.5 ---------------------------------------------------------------
// The tcgen05 kernels give one thread one pair row x one column "quarter" q of every fe layer, where
// quarter q is the set of 8-column chunks 4c + q (columns 32c + 8q + [0, 8), c = 0, 1, ...); layer 2 is
// handled as two halves of H2/2 columns with the same chunking inside each half.  Element i = 8c + e of
// a thread's slice is column 32c + 8q + e.  Draw 1 of (pair, quarter) -> words x,y: layer 1; word z: layer 2 low
// half; word w: high half.  Layer 0: for the default widths (H0 <= 96, H1 <= 160: edge_drop_compact) its 24 bits
// are bits 0..5 of every byte of word y -- layer 1's elements 32..39 only use bits 6,7 there -- i.e. the layer-0
// keep word is y << 2 and ONE Philox draw per (pair, quarter) serves all three layers (the draw is the largest
// single item of the dropout cost in the tcgen05 epilogues); wider layers take layer 0 from draw 0.  Inside a 32-bit word, element i
// uses bit edge_drop_bitpos(i & 31): the order in which byte-permutes with sign replication (PRMT)
// turn a word into packed bf16x2 / fp32 keep masks.  The generic kernel evaluates the same function
// element-wise.
__host__ __device__ __forceinline__ bool edge_drop_packed_ok(int H0, int H1, int H2) {
  return (H0 % 32 == 0) && (H1 % 32 == 0) && (H2 % 64 == 0) && H0 / 4 <= 128 && H1 / 4 <= 64 && H2 / 8 <= 32;
}
__host__ __device__ __forceinline__ u4 edge_drop_bits(uint64_t seed, uint64_t pair, uint32_t quarter, uint32_t draw) {
  return philox4x32_10((uint32_t)pair, (uint32_t)(pair >> 32), 0xED6E0000u | quarter, draw, (uint32_t)seed,
                       (uint32_t)(seed >> 32));
}
__host__ __device__ __forceinline__ bool edge_drop_compact(int H0, int H1) { return H0 <= 96 && H1 <= 160; }
__host__ __device__ __forceinline__ uint32_t edge_drop_bitpos(int e) {   // e in [0, 32)
  return 8u * (2u * ((uint32_t)(e >> 1) & 1u) + ((uint32_t)e & 1u)) + 7u - (uint32_t)(e >> 2);
}
__host__ __device__ __forceinline__ bool edge_drop_keep(uint64_t seed, uint64_t pair, int layer, int col, int H0,
                                                        int H1, int H2) {
  const int cc = layer == 2 ? col % (H2 / 2) : col;      // column inside the slice's tile (half)
  const int q = (cc >> 3) & 3;
  const int i = 8 * (cc >> 5) + (cc & 7);
  if (layer == 0 && edge_drop_compact(H0, H1)) {
    const u4 r = edge_drop_bits(seed, pair, (uint32_t)q, 1u);
    return (r.y >> (edge_drop_bitpos(i) - 2u)) & 1u;
  }
  const uint32_t draw = layer == 0 ? 0u : 1u;
  const int wsel = layer == 2 ? 2 + col / (H2 / 2) : (i >> 5);
  const u4 r = edge_drop_bits(seed, pair, (uint32_t)q, draw);
  const uint32_t word = wsel == 0 ? r.x : (wsel == 1 ? r.y : (wsel == 2 ? r.z : r.w));
  return (word >> edge_drop_bitpos(i & 31)) & 1u;
}

__device__ __forceinline__ float lrelu(float x, float a) { return x > 0.f ? x : a * x; }
__device__ __forceinline__ float lrelu_grad_from_out(float y, float a) { return y > 0.f ? 1.f : a; }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

static inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

}  // namespace mpg
