// Shared pieces of the tcgen05 / TMEM edge-network kernels (sm_100a): problem constants, shared-memory
// layout of the swizzled bf16 tiles, PTX wrappers (mbarrier, bulk-async copies, tcgen05 alloc / mma /
// commit / ld), UMMA descriptors and the weight-image kernel.  Included by edge_tc.cu inside its
// anonymous namespace.
#pragma once

constexpr int K0 = 96, N1 = 160, N2 = 192;
constexpr int TILE = 128;
constexpr int NQ = 4;                            // column quarters (warps sharing a TMEM lane group)
constexpr int Q0 = K0 / NQ, Q1 = N1 / NQ, Q2 = N2 / NQ;   // 24, 40, 48 columns per thread
constexpr int KSTEPS1 = K0 / 16 + 1;             // + bias step
constexpr int KSTEPS2 = N1 / 16 + 1;
constexpr uint32_t W1_BLK = N1 * 128;            // bytes of one 64-wide K block of W1 (rows = out features)
constexpr uint32_t W2_BLK = N2 * 128;
constexpr uint32_t A_BLK = TILE * 128;           // one 64-wide K block of an activation tile
constexpr uint32_t W1_BYTES = 2 * W1_BLK;        // K = 128 (96 + bias step, padded)
constexpr uint32_t W2_BYTES = 3 * W2_BLK;        // K = 192 (160 + bias step, padded)
constexpr uint32_t H0_BYTES = 2 * A_BLK;
constexpr uint32_t H1_BYTES = 3 * A_BLK;
constexpr uint32_t OFF_W1 = 0;
constexpr uint32_t OFF_W2 = OFF_W1 + W1_BYTES;   // 40960
constexpr uint32_t OFF_H0 = OFF_W2 + W2_BYTES;   // 114688
constexpr uint32_t OFF_H1 = OFF_H0 + H0_BYTES;   // 147456
constexpr uint32_t OFF_BAR = OFF_H1 + H1_BYTES;  // 196608
constexpr uint32_t SMEM_BYTES = OFF_BAR + 128 + 1024;   // + barriers + alignment slack
constexpr int NTHREADS = 128 * NQ;
constexpr uint32_t TMEM_COLS = 512, D1_COL = 0, D2_COL = 256;

// ---------------------------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// one arrival per warp (the barriers the epilogue warps signal are initialised with F_NEPI / 32 = 16): every lane has
// issued its own proxy / tcgen05 fence before; __syncwarp orders the lanes' accesses before the elected arrive
__device__ __forceinline__ void warp_arrive(uint32_t bar) {
  __syncwarp();
  if ((threadIdx.x & 31) == 0) mbar_arrive(bar);
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
#ifdef MPG_TEST_WAIT   // experiment: non-suspending poll
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
#else
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
#endif
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// bounded spin: a protocol bug must trap, never hang the GPU
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > (1u << 22)) {
#ifdef MPG_DEBUG_BARRIERS
      printf("mbarrier timeout: block %d thread %d barrier +%u parity %u\n", blockIdx.x, threadIdx.x, bar & 0xFFu, parity);
#endif
      __trap();
    }
  }
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(cols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T, bf16 operands, fp32 accumulate
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// K-major, 128-byte swizzle: 8-row groups 1024 B apart, descriptor version 1 (sm_100)
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
// the same descriptor split in two words: the high word is a constant, the low word carries the address
constexpr uint32_t UMMA_DESC_HI = 64u | (1u << 14) | (2u << 29);
__device__ __forceinline__ uint32_t umma_desc_lo(uint32_t saddr) { return ((saddr & 0x3FFFFu) >> 4) | (1u << 16); }
// D[tmem] (+)= A[tmem] * B[smem]^T: A is a bf16 [128 x 16] slice held in TMEM (lane = row, two K
// elements per 32-bit column => 8 columns per K = 16 step)
__device__ __forceinline__ void umma_bf16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// 4 / 8 consecutive 32-bit TMEM columns of this thread's lane
__device__ __forceinline__ void tmem_st4(uint32_t taddr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%2,%3,%4};" ::"r"(taddr), "r"(a), "r"(b), "r"(c), "r"(d)
               : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, uint32_t a, uint32_t b) {   // {a, b, b, b, b, b, b, b}
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%2,%2,%2,%2,%2,%2};" ::"r"(taddr), "r"(a), "r"(b)
               : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// Warp-collective forms: executed by ALL lanes of a converged warp; one elected lane (always the same
// one for a full mask) issues.  tcgen05.mma / commit / bulk copies take warp-uniform operands, and from a
// `lane == 0` branch ptxas wraps each of them in an elect-and-loop sequence that costs ~100 cycles per MMA.
__device__ __forceinline__ void umma_bf16_lo(uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p, e;\n\t.reg .b64 da, db;\n\t"
      "mov.b64 da, {%1, %5};\n\t"
      "mov.b64 db, {%2, %5};\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "@e tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, p;\n\t}"
      ::"r"(d_tmem), "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(accumulate), "r"(UMMA_DESC_HI)
      : "memory");
}
__device__ __forceinline__ void umma_commit_elect(uint32_t bar) {
  asm volatile(
      "{\n\t.reg .pred e;\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "@e tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}" ::"r"(bar)
      : "memory");
}
__device__ __forceinline__ void mbar_expect_tx_elect(uint32_t bar, uint32_t bytes) {
  asm volatile(
      "{\n\t.reg .pred e;\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "@e mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n\t}" ::"r"(bar), "r"(bytes)
      : "memory");
}
__device__ __forceinline__ void bulk_g2s_elect(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile(
      "{\n\t.reg .pred e;\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "@e cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n\t}" ::"r"(dst),
      "l"(src), "r"(bytes), "r"(bar)
      : "memory");
}
// CUTLASS-style leader election: ptxas knows the guarded region runs on exactly one lane, so its
// values are trivially warp-uniform and the single-thread instructions need no per-lane loop
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}
// one 16-byte reduction instead of four scalar atomics (addr 16-byte aligned)
__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
// hides a value from loop-invariant code motion
__device__ __forceinline__ void opaque(uint32_t& v) { asm volatile("" : "+r"(v)); }
__device__ __forceinline__ void opaque(uint64_t& v) { asm volatile("" : "+l"(v)); }
__host__ __device__ constexpr uint32_t umma_idesc(int N) {
  // c = f32 (1 << 4), a = b = bf16 (1 << 7, 1 << 10), K-major A and B, N >> 3 at bit 17, M >> 4 at bit 24
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(TILE >> 4) << 24);
}
// 16 / 8 consecutive fp32 columns of this thread's TMEM lane, into v[o..]
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float* v) {
  uint32_t r[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
// NC (multiple of 8) columns starting at taddr
template <int NC>
__device__ __forceinline__ void tmem_ld_cols(uint32_t taddr, float* v) {
#pragma unroll
  for (int c = 0; c + 16 <= NC; c += 16) tmem_ld16(taddr + c, v + c);
  if (NC % 16) tmem_ld8(taddr + (NC / 16) * 16, v + (NC / 16) * 16);
}

// Column ownership inside the tcgen05 kernels: thread (tile row, quarter q) owns the 8-column chunks
// 4c + q, c = 0, 1, ... of every tile / accumulator (columns 32c + 8q + [0, 8)).  A chunk is one 16-byte
// unit of the 128-byte-swizzled bf16 tiles, so all of a thread's tile stores are [x0 | x1] + constant.
// tmem_ld8xK: K such chunks (fp32) from TMEM with ONE wait; every destination register is an output of
// the asm statement that also holds the wait, so no consumer can be scheduled ahead of it.
__device__ __forceinline__ void tmem_ld8x5(uint32_t taddr, float* v) {
  uint32_t r[40];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%40];\n\t"
      "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%8,%9,%10,%11,%12,%13,%14,%15}, [%41];\n\t"
      "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%16,%17,%18,%19,%20,%21,%22,%23}, [%42];\n\t"
      "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%24,%25,%26,%27,%28,%29,%30,%31}, [%43];\n\t"
      "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%32,%33,%34,%35,%36,%37,%38,%39}, [%44];\n\t"
      "tcgen05.wait::ld.sync.aligned;"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31]), "=r"(r[32]), "=r"(r[33]), "=r"(r[34]), "=r"(r[35]), "=r"(r[36]), "=r"(r[37]), "=r"(r[38]), "=r"(r[39])
      : "r"(taddr + 0), "r"(taddr + 32), "r"(taddr + 64), "r"(taddr + 96), "r"(taddr + 128)
      : "memory");
#pragma unroll
  for (int i = 0; i < 40; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld8x3(uint32_t taddr, float* v) {
  uint32_t r[24];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%24];\n\t"
      "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%8,%9,%10,%11,%12,%13,%14,%15}, [%25];\n\t"
      "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%16,%17,%18,%19,%20,%21,%22,%23}, [%26];\n\t"
      "tcgen05.wait::ld.sync.aligned;"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23])
      : "r"(taddr + 0), "r"(taddr + 32), "r"(taddr + 64)
      : "memory");
#pragma unroll
  for (int i = 0; i < 24; ++i) v[i] = __uint_as_float(r[i]);
}

__device__ __forceinline__ void tmem_ld8x2(uint32_t taddr, float* v) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%16];\n\t"
      "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%8,%9,%10,%11,%12,%13,%14,%15}, [%17];\n\t"
      "tcgen05.wait::ld.sync.aligned;"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr + 0), "r"(taddr + 32)
      : "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 p = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&p);
}
// byte offset of the 16-byte chunk holding columns [k, k+8) of `row` inside a swizzled tile whose
// 64-wide K blocks are blk_bytes apart
__device__ __forceinline__ uint32_t swz_chunk(uint32_t row, uint32_t k, uint32_t blk_bytes) {
  const uint32_t blk = k >> 6, chunk = (k & 63) >> 3;
  return blk * blk_bytes + row * 128 + ((chunk ^ (row & 7)) << 4);
}
// store 8 consecutive columns (one 16-byte chunk) of a row as bf16
__device__ __forceinline__ void st_chunk(uint32_t addr, const float* v) {
  asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(pack_bf16(v[0], v[1])),
               "r"(pack_bf16(v[2], v[3])), "r"(pack_bf16(v[4], v[5])), "r"(pack_bf16(v[6], v[7])));
}
__device__ __forceinline__ void st_ones_chunk(uint32_t addr) {   // {1, 1, 0, 0, 0, 0, 0, 0}
  asm volatile("st.shared.v4.b32 [%0], {%1,%2,%2,%2};" ::"r"(addr), "r"(0x3F803F80u), "r"(0u));
}
__device__ __forceinline__ void st_zero_chunk(uint32_t addr) {
  asm volatile("st.shared.v4.b32 [%0], {%1,%1,%1,%1};" ::"r"(addr), "r"(0u));
}
// global[dst .. dst+bytes) += shared[src .. src+bytes) as fp32, asynchronously (bulk_group completion); 16-byte aligned
// addresses and size.  Issued by one thread.
__device__ __forceinline__ void bulk_reduce_add_f32(float* dst, uint32_t src_smem, uint32_t bytes) {
  asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;" ::"l"(dst), "r"(src_smem),
               "r"(bytes) : "memory");
}
// 32 contiguous bytes (32-byte aligned) of read-only global memory in one instruction: the epilogue threads read their
// rows with one row per lane, so every warp-wide load touches 32 different lines -- half as many instructions, half as
// many L1 wavefronts as two 16-byte loads
__device__ __forceinline__ void ldg256(const float* p, float (&v)[8]) {
  asm("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
      : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7])
      : "l"(p));
}
// keep bit `b` (0..127) of a 128-bit Philox draw as an all-ones / all-zeros word
__device__ __forceinline__ uint32_t keep_mask(const u4& bits, int b) {
  const uint32_t w = (b >> 5) == 0 ? bits.x : ((b >> 5) == 1 ? bits.y : ((b >> 5) == 2 ? bits.z : bits.w));
  int32_t m;
  asm("bfe.s32 %0, %1, %2, 1;" : "=r"(m) : "r"(w), "r"(b & 31));
  return (uint32_t)m;
}
__device__ __forceinline__ float apply_keep(float x, uint32_t mask) { return __uint_as_float(__float_as_uint(x) & mask); }

// Everything below is specific to the edge-network kernels; fn_tc.cu includes this header for the PTX wrappers
// above only (MPG_TC_WRAPPERS_ONLY).
#ifndef MPG_TC_WRAPPERS_ONLY
// ---- tile geometry: 128 consecutive padded rows, or (EdgeArgs::cmap) the rows listed by the compaction map -------------
__device__ __forceinline__ int tile_count(const EdgeArgs& a) {
  return a.cmap ? a.cmap[0] : (a.B * a.N + TILE - 1) / TILE;
}
// padded row (b*N + i) that lane `row` of `tile` stands for, -1 for none
__device__ __forceinline__ int tile_row(const EdgeArgs& a, int tile, int row) {
  if (a.cmap) return a.cmap[2 + 2 * a.ctiles_max + tile * TILE + row];
  const int r = tile * TILE + row;
  return r < a.B * a.N ? r : -1;
}
// first jet of the tile and the number of jets it spans
__device__ __forceinline__ void tile_jets(const EdgeArgs& a, int tile, int& j0, int& nj) {
  if (a.cmap) {
    j0 = a.cmap[2 + tile];
    nj = a.cmap[2 + a.ctiles_max + tile];
  } else {
    const int BN = a.B * a.N;
    j0 = (tile * TILE) / a.N;
    nj = min(tile * TILE + TILE - 1, BN - 1) / a.N - j0 + 1;
  }
}

struct TcArgs {
  EdgeArgs a;
  const uint8_t* w1img;   // pre-swizzled bf16 images (edge_setup_kernel: edge_prepare_block)
  const uint8_t* w2img;
  uint2* sbits;           // backward: sign bits of D2, one uint2 per (step, epilogue thread)
  float* wslab;           // backward: per-CTA weight-gradient partials (edge_tc_bwd.cuh: SLAB_*)
  // work list: the (tile, sender) steps with at least one unmasked sender row, tile-major, sender ascending
  // (step_list_kernel); CTA b owns steps [total*b/grid, total*(b+1)/grid)
  const int2* steps;
  const int* total_steps;
  int num_tiles;
};

// ---------------------------------------------------------------------------------------------------
// weight images: img[n][k] (K-major, SW128) = bf16(scale * W[n][k]) for k < K, bias_hi / bias_lo at
// k = K, K+1, zero elsewhere.
// ---------------------------------------------------------------------------------------------------
// One set-up launch per call: block 0 builds the work list, the next nb_pq blocks compute P / Q (pq_fwd_tile: job =
// 2 * tile + output), the remaining nb_prep blocks write both
// weight images and (forward) zero-fill the aggregate the edge kernel accumulates into with reductions.
struct PrepArgs {
  const float* W1; const float* b1; const float* W2; const float* b2;
  float scale;
  uint8_t* img1; uint8_t* img2;
  float4* agg4; size_t agg_n4;
};
__device__ __forceinline__ void edge_prepare_block(const PrepArgs& p, int blk, int nblk) {
  const int idx = blk * 256 + threadIdx.x;
  constexpr int E1 = N1 * 128, E2 = N2 * 192;
  if (idx < E1 + E2) {
    const bool first = idx < E1;
    const int j = first ? idx : idx - E1;
    const int Kpad = first ? 128 : 192, K = first ? K0 : N1, Nout = first ? N1 : N2;
    const float* W = first ? p.W1 : p.W2;
    const float* bias = first ? p.b1 : p.b2;
    uint8_t* img = first ? p.img1 : p.img2;
    const int n = j / Kpad, k = j % Kpad;
    float v = 0.f;
    if (k < K) v = W[(size_t)n * K + k] * p.scale;
    else if (k == K) v = bias[n];
    else if (k == K + 1) v = bias[n] - __bfloat162float(__float2bfloat16_rn(bias[n]));
    const uint32_t off = (uint32_t)(k >> 6) * (uint32_t)Nout * 128u + (uint32_t)n * 128u +
                         ((((uint32_t)(k & 63) >> 3) ^ ((uint32_t)n & 7u)) << 4) + (uint32_t)(k & 7) * 2u;
    *reinterpret_cast<__nv_bfloat16*>(img + off) = __float2bfloat16_rn(v);
  }
  const size_t stride = (size_t)nblk * 256;
  for (size_t i = idx; i < p.agg_n4; i += stride) p.agg4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
}

// ---------------------------------------------------------------------------------------------------
// Work list.  A step (128-receiver tile, sender index s) whose sender rows are masked in every jet the
// tile touches contributes exactly zero to the aggregate and to every gradient: it is dropped here, so
// padded particles cost nothing.  One block: a warp per tile counts its live senders (ballots), a scan over
// the tiles of a chunk gives every tile its slice, the warps write the slices (tile-major, sender ascending).
// ---------------------------------------------------------------------------------------------------
struct ListArgs {
  const float* mask; int B, N, num_tiles;
  int2* steps; int* total;
  int in_block;   // 1: the set-up kernel's last block builds the list; 0: step_list_kernel does (very large batches)
  // list slot i holds tile (i * stride) % num_tiles, stride coprime to num_tiles: batches arrive sorted by particle
  // count, and a CTA (a contiguous slice of the list) whose tiles all have few live senders would spend its time on
  // tile changes (~2 steps each) while the others wait -- interleaving long and short tiles balances the slices
  int stride;
  const int* cmap; int ctiles_max;   // receiver compaction (EdgeArgs): tile count and jets per tile come from the map
};
constexpr int LIST_CHUNK = 512;    // tiles per scan chunk
// shared memory of step_list_block: scan counters, the live-sender words of a chunk's tiles, one byte per four mask
// elements
__host__ __device__ inline size_t step_list_smem(long long BN, int N) {
  return (size_t)(LIST_CHUNK + 8) * sizeof(int) + (size_t)LIST_CHUNK * ((N + 31) / 32) * sizeof(uint32_t) +
         (size_t)((BN + 3) / 4 + 15) / 16 * 16;
}
__device__ __forceinline__ void step_list_block(const ListArgs& l, int* cnt /* shared, step_list_smem bytes */) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int BN = l.B * l.N, N = l.N, R = (N + 31) / 32;
  uint32_t* live = reinterpret_cast<uint32_t*>(cnt + LIST_CHUNK + 8);   // [LIST_CHUNK][R] live-sender bits of a tile
  uint8_t* nib = reinterpret_cast<uint8_t*>(live + LIST_CHUNK * R);     // bit (i & 3) of nib[i >> 2]: mask[i] != 0
  if (l.mask != nullptr) {   // one coalesced pass over the mask; everything below reads shared memory only
    const int n4 = BN >> 2;
    if ((reinterpret_cast<uintptr_t>(l.mask) & 15) == 0) {
      const float4* m4 = reinterpret_cast<const float4*>(l.mask);
#pragma unroll 4
      for (int i = threadIdx.x; i < n4; i += 256) {
        const float4 v = __ldg(m4 + i);
        nib[i] = (uint8_t)((v.x != 0.f) | ((v.y != 0.f) << 1) | ((v.z != 0.f) << 2) | ((v.w != 0.f) << 3));
      }
    } else {
      for (int i = threadIdx.x; i < n4; i += 256) {
        const float* m = l.mask + 4 * (size_t)i;
        nib[i] = (uint8_t)((m[0] != 0.f) | ((m[1] != 0.f) << 1) | ((m[2] != 0.f) << 2) | ((m[3] != 0.f) << 3));
      }
    }
    if (threadIdx.x == 0 && (BN & 3)) {
      uint8_t v = 0;
      for (int e = 0; e < (BN & 3); ++e) v |= (uint8_t)((l.mask[4 * (size_t)n4 + e] != 0.f) << e);
      nib[n4] = v;
    }
    __syncthreads();
  }
  const int num_tiles = l.cmap ? l.cmap[0] : l.num_tiles;
  int stride = l.stride;
  if (l.cmap) {   // the tile count is only known here: same choice of stride as the host makes
    stride = (int)(num_tiles * 0.381966f) | 1;
    auto gcd = [](int x, int y) { while (y) { const int r = x % y; x = y; y = r; } return x; };
    while (stride > 1 && gcd(stride, num_tiles) != 1) stride -= 2;
    if (stride < 1) stride = 1;
  }
  auto tile_of = [&](int slot) { return (int)(((long long)slot * stride) % num_tiles); };
  int base = 0;
  for (int t0 = 0; t0 < num_tiles; t0 += LIST_CHUNK) {
    const int nt = min(LIST_CHUNK, num_tiles - t0);
    for (int i = warp; i < nt; i += 8) {   // a warp per tile: live-sender words and their count
      const int tile = tile_of(t0 + i);
      int j0, j1;
      if (l.cmap) {
        j0 = l.cmap[2 + tile];
        j1 = j0 + l.cmap[2 + l.ctiles_max + tile] - 1;
      } else {
        j0 = (tile * TILE) / N;
        j1 = min(tile * TILE + TILE - 1, BN - 1) / N;
      }
      int count = 0;
      for (int r = 0; r < R; ++r) {
        const int s = 32 * r + lane;
        bool any = false;
        if (s < N) {
          any = l.mask == nullptr;
          if (!any)
            for (int j = j0; j <= j1; ++j) {
              const int e = j * N + s;
              any |= (nib[e >> 2] >> (e & 3)) & 1;
            }
        }
        const uint32_t b = __ballot_sync(0xffffffffu, any);
        if (lane == 0) live[i * R + r] = b;
        count += __popc(b);
      }
      if (lane == 0) cnt[i] = count;
    }
    __syncthreads();
    if (warp == 0) {   // exclusive scan of cnt[0, nt) in place, chunk total -> cnt[LIST_CHUNK]
      int run = 0;
      for (int i0 = 0; i0 < nt; i0 += 32) {
        const int v = i0 + lane < nt ? cnt[i0 + lane] : 0;
        int inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const int u = __shfl_up_sync(0xffffffffu, inc, o);
          if (lane >= o) inc += u;
        }
        if (i0 + lane < nt) cnt[i0 + lane] = run + inc - v;
        run += __shfl_sync(0xffffffffu, inc, 31);
      }
      if (lane == 0) cnt[LIST_CHUNK] = run;
    }
    __syncthreads();
    for (int i = warp; i < nt; i += 8) {   // the tile's slice: senders ascending
      int off = base + cnt[i];
      const int tile = tile_of(t0 + i);
      for (int r = 0; r < R; ++r) {
        const uint32_t b = live[i * R + r];
        if ((b >> lane) & 1u) l.steps[off + __popc(b & ((1u << lane) - 1u))] = make_int2(tile, 32 * r + lane);
        off += __popc(b);
      }
    }
    base += cnt[LIST_CHUNK];
    __syncthreads();
  }
  if (threadIdx.x == 0) *l.total = base;
}
// the same list from many blocks (one warp per tile, slices reserved with an atomic: tiles land in arbitrary order,
// only the grouping by tile matters to the kernels); *total must be zero on entry (the set-up kernel does it)
__global__ void __launch_bounds__(256) step_list_kernel(const float* __restrict__ mask, int B, int N, int num_tiles,
                                                        int2* __restrict__ steps, int* __restrict__ total,
                                                        const int* __restrict__ cmap, int ctiles_max) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tile = blockIdx.x * 8 + warp;
  if (tile >= (cmap ? cmap[0] : num_tiles)) return;
  const int BN = B * N;
  int j0, j1;
  if (cmap) {
    j0 = cmap[2 + tile];
    j1 = j0 + cmap[2 + ctiles_max + tile] - 1;
  } else {
    j0 = (tile * TILE) / N;
    j1 = min(tile * TILE + TILE - 1, BN - 1) / N;
  }
  auto active = [&](int s) {
    if (s >= N) return false;
    if (mask == nullptr) return true;
    bool any = false;
    for (int j = j0; j <= j1; ++j) any |= mask[(size_t)j * N + s] != 0.f;
    return any;
  };
  int count = 0;
  for (int s0 = 0; s0 < N; s0 += 32) count += __popc(__ballot_sync(0xffffffffu, active(s0 + lane)));
  int off = 0;
  if (lane == 0 && count > 0) off = atomicAdd(total, count);
  off = __shfl_sync(0xffffffffu, off, 0);
  for (int s0 = 0; s0 < N; s0 += 32) {
    const bool f = active(s0 + lane);
    const uint32_t b = __ballot_sync(0xffffffffu, f);
    if (f) steps[off + __popc(b & ((1u << lane) - 1u))] = make_int2(tile, s0 + lane);
    off += __popc(b);
  }
}

__global__ void __launch_bounds__(256) edge_setup_kernel(PqFwdArgs pq, int nb_pq, PrepArgs pr, int nb_prep, ListArgs ls) {
  extern __shared__ __align__(16) float setup_sm[];
  // the work-list block goes first: it is the longest single block (one SM walks every tile), so it must not wait for
  // a free slot behind hundreds of others
  const int nb_list = ls.in_block ? 1 : 0;
  const int b = (int)blockIdx.x - nb_list;
  if (b < 0) {
    step_list_block(ls, reinterpret_cast<int*>(setup_sm));
  } else if (b < nb_pq) {
    pq_fwd_tile(pq, b, setup_sm);
  } else {
    if (!ls.in_block && b == nb_pq && threadIdx.x == 0) *ls.total = 0;   // counter of the step_list_kernel that follows
    edge_prepare_block(pr, b - nb_pq, nb_prep);
  }
}
#endif  // MPG_TC_WRAPPERS_ONLY
