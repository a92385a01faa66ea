// tcgen05 backward of the edge network (included by edge_tc.cu inside its anonymous namespace).
//
// The backward recomputes H0'/H1'/D2 per (tile, sender) step exactly as the forward does and never
// stores an N^2 x hidden tensor.  All fp32 weight-gradient accumulators cannot be TMEM-resident next
// to the streaming tiles, so the work is split into two kernels over the SAME step partition:
//
//   CHAIN  H0' -> D1 -> H1' -> D2 -> G2' -> dH1 = G2' W2 -> G1' -> dH0 = G1' W1 -> G0'
//          dP[r] += G0' (registers); dW1^T += H0'^T G1' (TMEM accumulator, the constant-1 column of
//          the H0' tile makes row 96 of it db1); dQ[jet,s] = sum over the jet's rows of G0': one more
//          MMA, H0'^T G0', whose rows 98+j come from one-hot "row belongs to jet j" columns parked in
//          the unused columns of the H0' tile (no per-element atomics, no shuffles).
//          It also dumps the sign bits of D2 (one uint2 per thread and step) for the second kernel.
//   DW2    H0' -> D1 -> H1' (recompute of layer 1 only), G2' rebuilt from the dumped sign bits,
//          dW2 += G2'^T H1' in two M=128 TMEM accumulators; the constant-1 column of the H1' tile
//          (N = 176) makes column 160 of them db2.  No W2 image, no second-layer MMA.
//
// Scale conventions (sd = dropout scale 2 or 1, sl = (1+alpha)/2, os = out_scale, m = sender mask):
// activation tiles hold X' = X / (sd*sl) and the weight images sd*sl*W exactly as in the forward kernel,
// g(v) = v + cg|v| = lrelu(v)/sl.  Gradient tiles are multiplied only by the EXACT slope pair {1, alpha}
// (1 is exact in bf16, so only the small negative-slope terms carry bf16(alpha)'s 0.1% error):
//   G2+ = dAgg*m*keep2*{1,a}(D2) = G2/(sd*os)        dH1c = G2+ W2img = (sl/os) dH1
//   G1+ = dH1c*keep1*{1,a}(D1)  = (sl/(os*sd)) G1    dH0c = G1+ W1img = (sl^2/os) dH0
//   G0+ = dH0c*keep0*{1,a}(pre0) = (sl^2/(os*sd)) G0
//   dP,dQ = sd*os/sl^2 * G0+;  dW1 = sd^2*os * H0'^T G1+;  db1 = sd*os/sl * sum G1+;
//   dW2 = sd^2*sl*os * G2+^T H1';  db2 = sd*os * sum G2+.
// G2+/G1+ are produced in packed bf16x2 arithmetic (pack two fp32 values, multiply by a factor pair
// selected per lane from the sign bits: PRMT sign-replication mask + one LOP3); G0+ is formed in fp32 so
// that dP accumulates unrounded terms.  Sign bits are kept one word per 32 elements (sign_put / neg_*_mask).
//
// Structure as in the forward kernel: 16 epilogue warps (thread <-> tile row x column chunks 4c+q), a
// control warpgroup (warp 16 issues every MMA through an elected lane, warp 17 runs the TMA Q ring),
// setmaxnreg 112/32.  Operands that only feed an MMA as A live in TMEM (H1', G2', G1': written back in
// place over the fp32 accumulator they were computed from); tiles that feed an MMA as B or transposed
// (H0', G1', G0', H1', G2') are swizzled bf16 tiles in shared memory.  CHAIN is serial per step (TMEM
// holds one step) with H0'(s+1) built under M4(s), E4(s) under M1(s+1) and E1(s+1) under the dQ MMA of
// step s; in DW2 layer 1 runs one step ahead of the dW2 MMA, which keeps the tensor pipe busy.

enum { BWD_CHAIN = 0, BWD_DW2 = 1 };

// MN-major view of a SW128 tile: 64-element blocks along M/N are lbo_bytes apart, 8-row groups along
// K are 1024 B apart
__device__ __forceinline__ uint64_t umma_desc_mn(uint32_t saddr, uint32_t lbo_bytes) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16) | (64ull << 32) | (1ull << 46) |
         (2ull << 61);
}
__host__ __device__ constexpr uint32_t umma_idesc_t(int N, int a_mn, int b_mn) {
  return umma_idesc(N) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16);
}

// ---- shared-memory maps ------------------------------------------------------------------------------
// CHAIN: W1 | W2 | H0' (2 blocks) | X (3 blocks: G1', later G0') | Q ring (4 stages) | barriers
constexpr uint32_t C_OFF_H0 = OFF_W2 + W2_BYTES;            // 114688
constexpr uint32_t C_OFF_X = C_OFF_H0 + H0_BYTES;           // 147456
constexpr uint32_t C_OFF_Q = C_OFF_X + H1_BYTES;            // 196608
constexpr int C_QS = 4;
constexpr uint32_t C_OFF_BAR = C_OFF_Q + C_QS * F_QSTAGE;   // 211968
constexpr uint32_t C_SMEM = C_OFF_BAR + 256 + 1024;
// DW2: W1 | H0' | G2' (3 blocks; the 4th one the second M block reads is H1' buffer 0) | H1' x2 | Q ring (2)
constexpr uint32_t D_OFF_H0 = W1_BYTES;                     // 40960
constexpr uint32_t D_OFF_G2 = D_OFF_H0 + H0_BYTES;          // 73728
constexpr uint32_t D_OFF_H1 = D_OFF_G2 + H1_BYTES;          // 122880 (+ 49152 per buffer)
constexpr uint32_t D_OFF_Q = D_OFF_H1 + 2 * H1_BYTES;       // 221184
constexpr int D_QS = 2;
constexpr uint32_t D_OFF_BAR = D_OFF_Q + D_QS * F_QSTAGE;   // 228864
constexpr uint32_t D_SMEM = D_OFF_BAR + 256 + 1024;
// ---- TMEM maps -------------------------------------------------------------------------------------------
constexpr uint32_t C_RA = 0, C_RB = 160, C_PW1 = 352;       // CHAIN: D1/H1'/dH1/G1' | D2/G2'/dQ + dH0 | dW1^T
constexpr uint32_t D_R = 0, D_PW2A = 160, D_PW2B = 336;     // DW2: D1 | dW2 rows 0..127 | rows 128..191
constexpr int NDW = N1 + 16;                                 // 176: dW2 accumulator width incl. the db2 column
// weight-gradient partials: every CTA stores its TMEM accumulators to a private slab laid out like the
// final tensors (dW1 [160,96] | db1 [160] | dW2^T [160,192] | db2 [192]); wgrad_reduce_kernel sums the slabs.
// (148 CTAs x 46k atomics onto the same addresses cost more than the whole N=30 main loop.)
constexpr int SLAB_DW1 = 0, SLAB_DB1 = N1 * K0, SLAB_DW2 = SLAB_DB1 + N1, SLAB_DB2 = SLAB_DW2 + N2 * N1,
              SLAB_FLOATS = SLAB_DB2 + N2;

__device__ __forceinline__ uint32_t mul_bf16x2(uint32_t a, uint32_t b) {
  uint32_t d;
  asm("mul.rn.bf16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
  return d;
}
// sign bits of the bf16 pair `w` (bits 15 and 31) into slot p (0..15) of a 32-bit sign word:
// slots 0..7 use bits (15-p, 31-p), slots 8..15 bits (15-p, 31-p) of the low bytes, i.e. (7-(p-8), 23-(p-8))
__device__ __forceinline__ void sign_put(uint32_t& acc, uint32_t w, int p) {
  if (p < 8) acc |= (w >> p) & (0x80008000u >> p);
  else acc |= (w >> p) & (0x00800080u >> (p - 8));
}
// 0xFFFF per negative lane of pair p
__device__ __forceinline__ uint32_t neg_pair_mask(uint32_t sw, int p) {
  return p < 8 ? prmt(sw << p, 0u, 0xbb99u) : prmt(sw << (p - 8), 0u, 0xaa88u);
}
// all ones if lane `hi` (0 / 1) of pair p is negative
__device__ __forceinline__ uint32_t neg_lane_mask(uint32_t sw, int p, int hi) {
  return p < 8 ? prmt(sw << p, 0u, hi ? 0xbbbbu : 0x9999u) : prmt(sw << (p - 8), 0u, hi ? 0xaaaau : 0x8888u);
}
// per-lane select: negative lanes take `neg`, the others `pos`
__device__ __forceinline__ uint32_t sel_pair(uint32_t mask, uint32_t neg, uint32_t pos) {
  return (mask & neg) | (~mask & pos);
}
__device__ __forceinline__ uint32_t bf16x2_dup(float v) {
  const uint32_t h = (uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(v));
  return h | (h << 16);
}

struct BwdBars {   // byte offsets from the barrier block
  // rdy*: signalled by the 16 epilogue warps (one arrival each); done*: tcgen05.commit of the issuer.
  // CHAIN splits its hand-offs so that an MMA starts on the part of its operand that exists already:
  //   rdyB / rdyB1   H1' K columns [0,96) / [96,176) written (E1 rounds 0 / 1)   -> M2 K steps 0..5 / 6..10
  //   doneB / doneB1 D2 low / high N half complete                                  -> E2 half 0 / 1
  //   rdyC / rdyC1   G2' low / high half written                                    -> M3 K steps 0..5 / 6..11
  static constexpr uint32_t w = 0, rdyA = 8, rdyB = 16, rdyB1 = 24, rdyC = 32, rdyC1 = 40, rdyD = 48, rdyE = 56,
                            doneA = 64, doneB = 72, doneB1 = 80, doneC = 88, doneD5 = 96, doneD4 = 104, doneE = 112,
                            q = 128 /* [4] */, qe = 160 /* [4] */;
};

template <int MODE, bool DROP>
__global__ void __launch_bounds__(F_NTHR, 1) edge_tc_bwd_kernel(TcArgs t) {
  extern __shared__ uint8_t smem_raw[];
  const EdgeArgs& a = t.a;
  constexpr bool CH = MODE == BWD_CHAIN;
  constexpr uint32_t OFF_H0T = CH ? C_OFF_H0 : D_OFF_H0, OFF_QR = CH ? C_OFF_Q : D_OFF_Q,
                     OFF_BARS = CH ? C_OFF_BAR : D_OFF_BAR;
  constexpr int QS = CH ? C_QS : D_QS;
  uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* sm = smem_raw + (base - smem_u32(smem_raw));
  opaque(base);
  const uint32_t sQ = base + OFF_QR, bar0 = base + OFF_BARS;
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(sm + OFF_BARS + 192);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  MPG_TP(0);
  const long long total_steps = *t.total_steps;
  const long long g0 = total_steps * blockIdx.x / gridDim.x;
  const long long g1 = total_steps * (blockIdx.x + 1) / gridDim.x;
  const int nsteps = (int)(g1 - g0);
  const int2* steps = t.steps + g0;   // this CTA's (tile, sender) list
  const int N = a.N, BN = a.B * a.N;
  // every thread fetches the first record itself: the load is in flight under the barrier / TMEM set-up, and the
  // epilogue threads can start their first tile's row loads without waiting for the loader's first stage
  int2 first = make_int2(0, 0);
  if (nsteps > 0) first = __ldg(steps);

  if (threadIdx.x == 0) {
    mbar_init(bar0 + BwdBars::w, 1);
    for (uint32_t b = BwdBars::rdyA; b <= BwdBars::rdyE; b += 8) mbar_init(bar0 + b, F_NEPIW);
    for (uint32_t b = BwdBars::doneA; b <= BwdBars::doneE; b += 8) mbar_init(bar0 + b, 1);
    static_assert(BwdBars::rdyE + 8 == BwdBars::doneA && BwdBars::qe + 32 <= 192, "barrier block layout");
    for (int i = 0; i < 4; ++i) {
      mbar_init(bar0 + BwdBars::q + 8 * i, 1);
      mbar_init(bar0 + BwdBars::qe + 8 * i, F_NEPIW);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 16) tmem_alloc(smem_u32(const_cast<uint32_t*>(tmem_slot)), TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  MPG_TP(1);

  if (warp >= 16) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(F_REGS_CTL));
    const uint32_t ub = (smem_u32(smem_raw) + 1023u) & ~1023u;   // warp-uniform copy of `base`
    const uint32_t ubar = ub + OFF_BARS;
    if (warp == 17 && nsteps > 0) {
      // =============================== TMA loader: weight images, Q ring =============================
      if (CH) {
        mbar_expect_tx_elect(ubar + BwdBars::w, W1_BYTES + W2_BYTES);
        bulk_g2s_elect(ub + OFF_W1, t.w1img, W1_BYTES, ubar + BwdBars::w);
        bulk_g2s_elect(ub + OFF_W2, t.w2img, W2_BYTES, ubar + BwdBars::w);
      } else {
        mbar_expect_tx_elect(ubar + BwdBars::w, W1_BYTES);
        bulk_g2s_elect(ub + OFF_W1, t.w1img, W1_BYTES, ubar + BwdBars::w);
      }
      int2 ts = first;
      int l_tile = -1, j0 = 0, nj = 1;   // jets of the tile whose steps are being loaded
      for (int it = 0; it < nsteps; ++it) {
        const int2 ts_next = steps[it + 1 < nsteps ? it + 1 : it];   // in flight while this step's copies are issued
        const int st = it % QS;
        if (it >= QS) mbar_wait(ubar + BwdBars::qe + 8 * st, (it / QS - 1) & 1);
        const int q_tile = ts.x, q_s = ts.y;
        if (q_tile != l_tile) { tile_jets(a, q_tile, j0, nj); l_tile = q_tile; }
        const uint32_t bar = ubar + BwdBars::q + 8 * st;
        const uint32_t dst = ub + OFF_QR + (uint32_t)st * F_QSTAGE;
        // stage header (tile, sender, the sender's mask in each jet of the tile): see edge_tc_fwd.cuh
        // the mask load and the row copies are in flight together; the header is stored and the barrier's own
        // arrival (with the byte count) comes last, so the phase cannot complete before the header is visible
        float mv = 1.f;
        if (lane < nj && a.mask) mv = __ldg(a.mask + (size_t)(j0 + lane) * N + q_s);
        for (int j = 0; j < nj; ++j)
          bulk_g2s_elect(dst + (uint32_t)j * F_QROW, a.Q + ((size_t)(j0 + j) * N + q_s) * K0, F_QROW, bar);
        if (lane < nj) asm volatile("st.shared.f32 [%0], %1;" ::"r"(dst + F_QHDR + 8 + 4 * (uint32_t)lane), "f"(mv) : "memory");
        if (lane == 0) asm volatile("st.shared.v2.b32 [%0], {%1,%2};" ::"r"(dst + F_QHDR), "r"(q_tile), "r"(q_s) : "memory");
        __syncwarp();
        mbar_expect_tx_elect(bar, (uint32_t)nj * F_QROW);
        ts = ts_next;
      }
    } else if (warp == 16 && nsteps > 0) {
      // =============================== MMA issuer ======================================================
      uint64_t dH0 = umma_desc(ub + OFF_H0T), dW1 = umma_desc(ub + OFF_W1);
      auto m1 = [&]() {   // D1 = H0' W1'^T  -> CHAIN: RA, DW2: R (both column 0)
        if (elect_one()) {
          opaque(dH0); opaque(dW1);
#pragma unroll
          for (uint32_t ks = 0; ks < KSTEPS1; ++ks) {
            const uint32_t blk = ks >> 2, j = ks & 3;
            umma_bf16(tmem, dH0 + ((blk * A_BLK + j * 32) >> 4), dW1 + ((blk * W1_BLK + j * 32) >> 4),
                      umma_idesc(N1), ks);
          }
          umma_commit(ubar + BwdBars::doneA);
        }
        __syncwarp();
      };
      mbar_wait(ubar + BwdBars::w, 0);
      mbar_wait(ubar + BwdBars::rdyA, 0);
      tc_fence_after();
      m1();
      if constexpr (CH) {
        uint64_t dW2 = umma_desc(ub + OFF_W2);
        for (int it = 0; it < nsteps; ++it) {
          const uint32_t par = it & 1;
          // ---- M2: D2 = H1'(TMEM RA) W2'^T -> RB as two N = 96 halves.  K steps 0..5 read H1' columns [0,96), which
          // E1's first round has written (rdyB); the rest follows after the second round (rdyB1).  The low half is
          // committed first so that E2 starts on it while the high half is still in the tensor pipe.
          mbar_wait(ubar + BwdBars::rdyB, par);
          MPG_TR(it, 11);
          tc_fence_after();
          if (elect_one()) {
            opaque(dW2);
#pragma unroll
            for (uint32_t ks = 0; ks < 6; ++ks) {
              const uint32_t blk = ks >> 2, j = ks & 3;
              umma_bf16_ts(tmem + C_RB, tmem + C_RA + ks * 8, dW2 + ((blk * W2_BLK + j * 32) >> 4), umma_idesc(NH2), ks);
              umma_bf16_ts(tmem + C_RB + NH2, tmem + C_RA + ks * 8, dW2 + ((blk * W2_BLK + NH2 * 128 + j * 32) >> 4),
                           umma_idesc(NH2), ks);
            }
          }
          __syncwarp();
          mbar_wait(ubar + BwdBars::rdyB1, par);
          tc_fence_after();
          if (elect_one()) {
            opaque(dW2);
#pragma unroll
            for (uint32_t ks = 6; ks < KSTEPS2; ++ks) {
              const uint32_t blk = ks >> 2, j = ks & 3;
              umma_bf16_ts(tmem + C_RB, tmem + C_RA + ks * 8, dW2 + ((blk * W2_BLK + j * 32) >> 4), umma_idesc(NH2), 1u);
            }
            umma_commit(ubar + BwdBars::doneB);
#pragma unroll
            for (uint32_t ks = 6; ks < KSTEPS2; ++ks) {
              const uint32_t blk = ks >> 2, j = ks & 3;
              umma_bf16_ts(tmem + C_RB + NH2, tmem + C_RA + ks * 8, dW2 + ((blk * W2_BLK + NH2 * 128 + j * 32) >> 4),
                           umma_idesc(NH2), 1u);
            }
            umma_commit(ubar + BwdBars::doneB1);
          }
          __syncwarp();
          MPG_TR(it, 12);
          // ---- M3: dH1 = G2'(TMEM RB) W2' (B = W2 image read MN-major: N = 160 inputs, K = 192 outputs) -> RA.
          // K steps 0..5 take the low half of G2' (rdyC), 6..11 the high half (rdyC1).  RA still holds H1' until M2's
          // last K step has executed: the tensor pipe runs the MMAs of one CTA in issue order.
          mbar_wait(ubar + BwdBars::rdyC, par);
          MPG_TR(it, 13);
          tc_fence_after();
          if (elect_one()) {
            uint64_t dB = umma_desc_mn(ub + OFF_W2, W2_BLK);
            opaque(dB);
#pragma unroll
            for (uint32_t ks = 0; ks < 6; ++ks)
              umma_bf16_ts(tmem + C_RA, tmem + C_RB + ks * 8, dB + ((ks * 2048) >> 4), umma_idesc_t(N1, 0, 1), ks);
          }
          __syncwarp();
          mbar_wait(ubar + BwdBars::rdyC1, par);
          tc_fence_after();
          if (elect_one()) {
            uint64_t dB = umma_desc_mn(ub + OFF_W2, W2_BLK);
            opaque(dB);
#pragma unroll
            for (uint32_t ks = 6; ks < N2 / 16; ++ks)
              umma_bf16_ts(tmem + C_RA, tmem + C_RB + ks * 8, dB + ((ks * 2048) >> 4), umma_idesc_t(N1, 0, 1), 1u);
            umma_commit(ubar + BwdBars::doneC);
          }
          __syncwarp();
          // ---- M5: dW1^T += H0'^T G1' (both MN-major from shared memory); M4: dH0 = G1'(TMEM RA) W1' -> RB+96
          mbar_wait(ubar + BwdBars::rdyD, par);
          MPG_TR(it, 14);
          tc_fence_after();
          if (elect_one()) {
            uint64_t dA = umma_desc_mn(ub + C_OFF_H0, A_BLK), dB = umma_desc_mn(ub + C_OFF_X, A_BLK);
            opaque(dA); opaque(dB);
#pragma unroll
            for (uint32_t ks = 0; ks < TILE / 16; ++ks)
              umma_bf16(tmem + C_PW1, dA + ((ks * 2048) >> 4), dB + ((ks * 2048) >> 4), umma_idesc_t(N1, 1, 1),
                        (it > 0 || ks > 0) ? 1u : 0u);
            umma_commit(ubar + BwdBars::doneD5);
            uint64_t dW = umma_desc_mn(ub + OFF_W1, W1_BLK);
            opaque(dW);
#pragma unroll
            for (uint32_t ks = 0; ks < N1 / 16; ++ks)
              umma_bf16_ts(tmem + C_RB + NH2, tmem + C_RA + ks * 8, dW + ((ks * 2048) >> 4), umma_idesc_t(K0, 0, 1), ks);
            umma_commit(ubar + BwdBars::doneD4);
          }
          __syncwarp();
          // ---- M1 of the next step (its H0' tile is built under M4) ---------------------------------------------
          if (it + 1 < nsteps) {
            mbar_wait(ubar + BwdBars::rdyA, (it + 1) & 1);
            tc_fence_after();
            m1();
          }
          // ---- M7: per-jet column sums of G0' = rows 98+j of H0'^T G0' -> RB[0,96) --------------------------------
          mbar_wait(ubar + BwdBars::rdyE, par);
          MPG_TR(it, 15);
          tc_fence_after();
          if (elect_one()) {
            uint64_t dA = umma_desc_mn(ub + C_OFF_H0, A_BLK), dB = umma_desc_mn(ub + C_OFF_X, A_BLK);
            opaque(dA); opaque(dB);
#pragma unroll
            for (uint32_t ks = 0; ks < TILE / 16; ++ks)
              umma_bf16(tmem + C_RB, dA + ((ks * 2048) >> 4), dB + ((ks * 2048) >> 4), umma_idesc_t(K0, 1, 1), ks);
            umma_commit(ubar + BwdBars::doneE);
          }
          __syncwarp();
        }
      } else {
        for (int it = 0; it < nsteps; ++it) {
          const uint32_t par = it & 1;
          mbar_wait(ubar + BwdBars::rdyB, par);          // E1(it) done: H1'(it) in shared memory, D1 free
          if (it + 1 < nsteps) {                          // layer 1 runs one step ahead
            mbar_wait(ubar + BwdBars::rdyA, (it + 1) & 1);
            tc_fence_after();
            m1();
          }
          // ---- M6: dW2 += G2'^T H1' (two M blocks over the G2' columns, N = 176, K = 128 rows) ----------------------
          mbar_wait(ubar + BwdBars::rdyC, par);          // G2'(it) built
          tc_fence_after();
          if (elect_one()) {
            uint64_t dA = umma_desc_mn(ub + D_OFF_G2, A_BLK), dB = umma_desc_mn(ub + D_OFF_H1 + par * H1_BYTES, A_BLK);
            opaque(dA); opaque(dB);
#pragma unroll
            for (uint32_t mb = 0; mb < 2; ++mb)
#pragma unroll
              for (uint32_t ks = 0; ks < TILE / 16; ++ks)
                umma_bf16(tmem + (mb ? D_PW2B : D_PW2A), dA + ((mb * 2 * A_BLK + ks * 2048) >> 4),
                          dB + ((ks * 2048) >> 4), umma_idesc_t(NDW, 1, 1), (it > 0 || ks > 0) ? 1u : 0u);
            umma_commit(ubar + BwdBars::doneC);
          }
          __syncwarp();
        }
      }
    }
  } else {
    // =============================== epilogue warps ====================================================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(F_REGS_EPI));
    MPG_TP(8);
    if (nsteps == 0) {   // no work (fewer live steps than CTAs): this CTA's part of its slab must still be defined
      float* slab = t.wslab + (size_t)blockIdx.x * SLAB_FLOATS;
      for (int i = (CH ? SLAB_DW1 : SLAB_DW2) + threadIdx.x; i < (CH ? SLAB_DW2 : SLAB_FLOATS); i += F_NEPI) slab[i] = 0.f;
    }
    if (nsteps > 0) {
      const int q = warp >> 2;
      const int row = (warp & 3) * 32 + lane;
      const uint32_t tl = tmem + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)q * 8;   // chunk q of this lane
      const uint32_t tp = tl - (uint32_t)q * 4;                                            // packed column 4q
      const float cg = (1.f - a.alpha) / (1.f + a.alpha);
      const float sd = DROP ? 2.f : 1.f, sl = 0.5f * (1.f + a.alpha), os = a.out_scale;
      const float sc_g0 = sd * os / (sl * sl), sc_dw1 = sd * sd * os, sc_db1 = sd * os / sl, sc_dw2 = sd * sd * sl * os,
                  sc_db2 = sd * os;
      DropCfg drop = a.drop;
      if (DROP) resolve_seed(drop);
      const uint32_t F_NEG = bf16x2_dup(a.alpha), F_POS = 0x3F803F80u;   // slope pair {alpha, 1}
      // this thread's two 16-byte positions inside a swizzled 128-byte tile row (chunk 4c+q -> xs{c&1} + (c>>1)*A_BLK)
      uint32_t xs0 = base + (uint32_t)row * 128u + ((uint32_t)(q ^ (row & 7)) << 4);
      uint32_t xs1 = base + (uint32_t)row * 128u + ((uint32_t)((4 + q) ^ (row & 7)) << 4);
      opaque(xs0);
      opaque(xs1);
      const uint32_t sH0 = base + OFF_H0T;
      const int bar_id = 1 + (warp & 3);      // named barrier of the four warps sharing this TMEM lane quarter

      float Preg[Q0];
      uint32_t dAggp[Q2 / 2];                 // bf16x2 pairs of dAgg[r][h*96 + 32c + 8q + ..]
      float dPacc[CH ? Q0 : 1];
      uint32_t s0 = 0, k0w = 0;               // CHAIN: pre0 sign word / layer-0 keep word of the current step
      uint32_t s0_next = 0, k0w_next = 0;     // ... of the step whose H0' was built last

      // ---- H0' builder state -------------------------------------------------------------------------------
      int h_loaded = -1, h_r = 0;
      uint32_t h_qoff = 0, h_moff = 0;
      bool h_valid = false;
      // record of the step whose H0' was built last (tile, sender, this row's mask multiplier)
      int n_tile = -1, n_s = 0;
      float n_m = 0.f;
      u4 n_kb{0, 0, 0, 0};                    // DROP: that step's Philox draw (layer 0 = y << 2; x,y: layer 1; z,w: layer 2)
      auto write_onehot = [&](int tile) {   // CHAIN: constant-1 columns 96,97 and one-hot jet columns 98+j of the H0' tile
        if (q == 0) {
          int tj0, tnj;
          tile_jets(a, tile, tj0, tnj);
          const int r = tile_row(a, tile, row);
          const int js = r >= 0 ? r / N - tj0 : 0;   // 0 .. F_QJ-1
          uint32_t w[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) w[i] = 0u;
          w[0] = 0x3F803F80u;
          if (r >= 0) {
#pragma unroll
            for (int i = 1; i < 8; ++i)
              if (i == ((2 + js) >> 1)) w[i] = 0x3F80u << (16 * (js & 1));
          }
          asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(sH0 + swz_chunk(row, 96, A_BLK)), "r"(w[0]),
                       "r"(w[1]), "r"(w[2]), "r"(w[3]));
          asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(sH0 + swz_chunk(row, 104, A_BLK)), "r"(w[4]),
                       "r"(w[5]), "r"(w[6]), "r"(w[7]));
          st_zero_chunk(sH0 + swz_chunk(row, 112, A_BLK));
          st_zero_chunk(sH0 + swz_chunk(row, 120, A_BLK));
        }
      };
      auto tile_rows = [&](int tile) {   // this thread's row of `tile`: index, stage offsets, its P values
        h_loaded = tile;
        int tj0, tnj;
        tile_jets(a, tile, tj0, tnj);
        const int r = tile_row(a, tile, row);               // padded row of this lane, -1: none
        h_valid = r >= 0;
        // lanes without a row compute on a row that exists (their mask multiplier is 0)
        const int rc = h_valid ? r : (a.cmap ? tj0 * N : BN - 1);
        h_r = rc;
        const uint32_t jl = (uint32_t)(rc / N - tj0);
        h_moff = F_QHDR + 8 + 4 * jl;
        h_qoff = jl * F_QROW + (uint32_t)q * 32u;
        // row-major: 8 floats at column 32c + 8q; tiled (EdgeArgs::p_tiled): column groups 8c + 2q (+1), the warp's 32
        // rows of one group contiguous, indexed by the lane's position in the tile
        const int prow = (a.cmap || h_valid) ? row : rc - tile * TILE;
        const float* p = a.p_tiled ? a.P + (size_t)tile * TILE * K0 + ((size_t)(2 * q) * TILE + prow) * 4
                                   : a.P + (size_t)rc * K0 + q * 8;
        const int cs = a.p_tiled ? 8 * TILE * 4 : 32, hs = a.p_tiled ? TILE * 4 : 4;
#pragma unroll
        for (int c = 0; c < Q0 / 8; ++c) {
          const float4 v0 = __ldg(reinterpret_cast<const float4*>(p + cs * c));
          const float4 v1 = __ldg(reinterpret_cast<const float4*>(p + cs * c + hs));
          Preg[8 * c] = v0.x; Preg[8 * c + 1] = v0.y; Preg[8 * c + 2] = v0.z; Preg[8 * c + 3] = v0.w;
          Preg[8 * c + 4] = v1.x; Preg[8 * c + 5] = v1.y; Preg[8 * c + 6] = v1.z; Preg[8 * c + 7] = v1.w;
        }
      };
      auto build_h0 = [&](int it) {
        mbar_wait(bar0 + BwdBars::q + 8 * (it % QS), (it / QS) & 1);
        if (it == 0) MPG_TP(12);
        const uint32_t stage = sQ + (uint32_t)(it % QS) * F_QSTAGE;
        int h_tile, h_s;
        asm volatile("ld.shared.v2.b32 {%0,%1}, [%2];" : "=r"(h_tile), "=r"(h_s) : "r"(stage + F_QHDR));
        if (h_tile != h_loaded) tile_rows(h_tile);
        {
          float mv;
          asm volatile("ld.shared.f32 %0, [%1];" : "=f"(mv) : "r"(stage + h_moff));
          n_tile = h_tile; n_s = h_s; n_m = h_valid ? mv : 0.f;
        }
        if (DROP) {
          n_kb = edge_drop_bits(drop.seed, (uint64_t)h_r * N + h_s, q, 1);
          k0w_next = n_kb.y << 2;
        }
        const uint32_t qa = stage + h_qoff;
        s0_next = 0;
#pragma unroll
        for (int c = 0; c < Q0 / 8; ++c) {
          float v[8];
          asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];"
                       : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]) : "r"(qa + c * 128));
          asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];"
                       : "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7]) : "r"(qa + c * 128 + 16));
          uint32_t w[4];
#pragma unroll
          for (int e = 0; e < 8; e += 2) {
            w[e >> 1] = pack_bf16(lrelu_g(v[e] + Preg[8 * c + e], cg), lrelu_g(v[e + 1] + Preg[8 * c + e + 1], cg));
            if (CH) sign_put(s0_next, w[e >> 1], 4 * c + (e >> 1));
            if (DROP) w[e >> 1] &= keep_pair(k0w_next, 8 * c + e);
          }
          asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(((c & 1) ? xs1 : xs0) + OFF_H0T + (c >> 1) * A_BLK),
                       "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]));
        }
        fence_async_smem();
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(bar0 + BwdBars::qe + 8 * (it % QS));
          mbar_arrive(bar0 + BwdBars::rdyA);
        }
      };

      // ---- state of the current step ------------------------------------------------------------------------
      int c_tile = 0, c_s = 0, c_r = 0, c_j0 = 0, c_nj = 1;   // current step's tile, sender, this lane's row, the tile's jets
      float c_m = 0.f;                        // mask multiplier of the current step (0 for rows past the end)
      bool c_valid = false, c_first = true;   // c_first: first step of its tile inside this CTA's range
      // DW2 builds two steps ahead: records of steps it+1 (p1_*) and it+2 (n_*)
      int p1_tile = -1, p1_s = 0;
      float p1_m = 0.f;
      u4 kb{0, 0, 0, 0}, p1_kb{0, 0, 0, 0};   // draws of the current step / of step it+1 (DW2)
      const bool dagg32 = (reinterpret_cast<uintptr_t>(a.dagg) & 31) == 0;
      auto enter_tile = [&]() {   // rows of the tile the current step belongs to; their dAgg
        const int r = tile_row(a, c_tile, row);
        tile_jets(a, c_tile, c_j0, c_nj);
        c_valid = r >= 0;
        c_r = c_valid ? r : (a.cmap ? c_j0 * N : BN - 1);
        const float* dg = a.dagg + (size_t)c_r * N2 + q * 8;
        if (dagg32) {   // 32-byte loads (see ldg256)
#pragma unroll
          for (int h = 0; h < 2; ++h)
#pragma unroll
            for (int c = 0; c < QH / 8; ++c) {
              float v[8];
              ldg256(dg + h * NH2 + 32 * c, v);
#pragma unroll
              for (int e = 0; e < 4; ++e) dAggp[h * (QH / 2) + 4 * c + e] = pack_bf16(v[2 * e], v[2 * e + 1]);
            }
        } else {
#pragma unroll
          for (int h = 0; h < 2; ++h)
#pragma unroll
            for (int c = 0; c < QH / 8; ++c) {
              const float4 v0 = __ldg(reinterpret_cast<const float4*>(dg + h * NH2 + 32 * c));
              const float4 v1 = __ldg(reinterpret_cast<const float4*>(dg + h * NH2 + 32 * c + 4));
              dAggp[h * (QH / 2) + 4 * c] = pack_bf16(v0.x, v0.y);
              dAggp[h * (QH / 2) + 4 * c + 1] = pack_bf16(v0.z, v0.w);
              dAggp[h * (QH / 2) + 4 * c + 2] = pack_bf16(v1.x, v1.y);
              dAggp[h * (QH / 2) + 4 * c + 3] = pack_bf16(v1.z, v1.w);
            }
        }
      };
      if (CH) {
#pragma unroll
        for (int c = 0; c < Q0; ++c) dPacc[c] = 0.f;
      } else {
        if (q == 0) {
          st_ones_chunk(sH0 + swz_chunk(row, 96, A_BLK));
          st_zero_chunk(sH0 + swz_chunk(row, 104, A_BLK));
        }
        if (q == 1) {   // constant-1 columns 160,161 of both H1' buffers (db2 column of the dW2 accumulators)
#pragma unroll
          for (int b = 0; b < 2; ++b) {
            st_ones_chunk(base + D_OFF_H1 + b * H1_BYTES + swz_chunk(row, 160, A_BLK));
            st_zero_chunk(base + D_OFF_H1 + b * H1_BYTES + swz_chunk(row, 168, A_BLK));
          }
        }
      }
      // the first tile's P and dAgg rows are requested before the loader's first stage is waited for (one memory
      // round trip under the other)
      MPG_TP(9);
      tile_rows(first.x);
      c_tile = first.x;
      enter_tile();
      MPG_TP(10);
      if (CH) write_onehot(first.x);   // constant-1 / one-hot columns: in place before build_h0(0) releases M1(0)
      MPG_TP(11);
      build_h0(0);
      MPG_TP(13);
      c_tile = n_tile; c_s = n_s; c_m = n_m;
      kb = n_kb;
      s0 = s0_next;
      k0w = k0w_next;
      if (!CH && nsteps > 1) {   // DW2: layer 1 runs one step ahead
        mbar_wait(bar0 + BwdBars::doneA, 0);
        build_h0(1);
        p1_tile = n_tile; p1_s = n_s; p1_m = n_m; p1_kb = n_kb;
      }

      int p_j0 = 0, p_nj = 0, p_s = 0;   // CHAIN: the step whose per-jet dQ sums sit in RB[0,96)
      auto dq_readout = [&]() {          // rows 98+j of RB[0,96) -> dQ[(j0+j)*N + s]: a handful of lanes
        if ((warp & 3) == 3) {
          float v[QH];
          tmem_ld8x3(tl + C_RB, v);
          if (lane >= 2 && lane < 2 + p_nj) {
            float* dq = a.dQ + ((size_t)(p_j0 + lane - 2) * N + p_s) * K0 + q * 8;
#pragma unroll
            for (int e = 0; e < QH; e += 4)
              red_add_v4(dq + 32 * (e >> 3) + (e & 7), v[e] * sc_g0, v[e + 1] * sc_g0, v[e + 2] * sc_g0, v[e + 3] * sc_g0);
          }
        }
      };

      MPG_TP(2);
      for (int it = 0; it < nsteps; ++it) {
        const uint32_t par = it & 1;
        const float mfac = c_m;

        // ---- E1: D1 -> H1' (+ sign words): chunks 0..2, then 3..4 (24 / 16 live values) -----------------------------
        uint32_t s1[2] = {0, 0};
        MPG_TRW(it, 0);
        mbar_wait(bar0 + BwdBars::doneA, par);
        MPG_TRW(it, 1);
        tc_fence_after();
        auto e1_round = [&](auto RND) {
          constexpr int rnd = decltype(RND)::value, nc = rnd == 0 ? 3 : 2, c0 = rnd * 3;
          float v[8 * nc];
          if constexpr (rnd == 0) tmem_ld8x3(tl + (CH ? C_RA : D_R), v);
          else tmem_ld8x2(tl + (CH ? C_RA : D_R) + 96, v);
          uint32_t w[4 * nc];
#pragma unroll
          for (int i = 0; i < 8 * nc; i += 2) {
            const int el = 8 * c0 + i, p = el >> 1;   // element / pair index inside the thread's 40-wide slice
            w[i >> 1] = pack_bf16(lrelu_g(v[i], cg), lrelu_g(v[i + 1], cg));
            if (CH) sign_put(s1[p >> 4], w[i >> 1], p & 15);
            if (DROP) w[i >> 1] &= keep_pair(el < 32 ? kb.x : kb.y, el & 31);
          }
          if constexpr (CH) {
            // in place: packed columns 16c+4q.. overlap fp32 columns other warps of this lane quarter read
            // in round 0 (chunks 0..2); by round 1 everything below column 96 has been consumed
            if constexpr (rnd == 0) named_bar_sync(bar_id, 128);
#pragma unroll
            for (int c = 0; c < nc; ++c)
              tmem_st4(tp + C_RA + 16 * (c0 + c), w[4 * c], w[4 * c + 1], w[4 * c + 2], w[4 * c + 3]);
          } else {
#pragma unroll
            for (int c = 0; c < nc; ++c) {
              const int cc = c0 + c;
              asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(((cc & 1) ? xs1 : xs0) + D_OFF_H1 + par * H1_BYTES + (cc >> 1) * A_BLK),
                           "r"(w[4 * c]), "r"(w[4 * c + 1]), "r"(w[4 * c + 2]), "r"(w[4 * c + 3]));
            }
          }
        };
        e1_round(std::integral_constant<int, 0>{});
        if constexpr (CH) {
          tmem_st_wait();
          // dQ of the previous step must leave RB before M2 of this step overwrites it
          if (it >= 1) {
            mbar_wait(bar0 + BwdBars::doneE, (it - 1) & 1);
            tc_fence_after();
            dq_readout();
            if (c_first) write_onehot(c_tile);   // first step of a new tile: the old one-hot columns are free now
          }
          tc_fence_before();
          warp_arrive(bar0 + BwdBars::rdyB);     // H1' columns [0,96) in place: M2's K steps 0..5 may start
        }
        e1_round(std::integral_constant<int, 1>{});
        if constexpr (CH) {
          if (q == 0) tmem_st8(tl + C_RA + N1 / 2, 0x3F803F80u, 0u);   // bias K-step: columns 160,161 = 1.0
          tmem_st_wait();
          tc_fence_before();
          warp_arrive(bar0 + BwdBars::rdyB1);
        } else {
          fence_async_smem();
          tc_fence_before();
          warp_arrive(bar0 + BwdBars::rdyB);
        }
        MPG_TRW(it, 2);

        // ---- G2' = dAgg * m * keep2 * (1 + cg sgn D2) ----------------------------------------------------------------
        const uint32_t U_POS = bf16x2_dup(mfac), U_NEG = bf16x2_dup(mfac * a.alpha);
        if constexpr (CH) {
          uint2 sb{0, 0};
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            mbar_wait(bar0 + (h ? BwdBars::doneB1 : BwdBars::doneB), par);   // this N half of D2 is complete
            if (h == 0) MPG_TRW(it, 3);
            tc_fence_after();
            float v[QH];
            tmem_ld8x3(tl + C_RB + h * NH2, v);
            uint32_t sw = 0, gw[QH / 2];
#pragma unroll
            for (int i = 0; i < QH; i += 2) {
              const int p = i >> 1;
              const uint32_t zw = pack_bf16(v[i], v[i + 1]);
              sign_put(sw, zw, p);
#ifdef MPG_FP32_SLOPE
              {
                const uint32_t dw = dAggp[h * (QH / 2) + p];
                const float f0 = v[i] < 0.f ? mfac * a.alpha : mfac, f1 = v[i + 1] < 0.f ? mfac * a.alpha : mfac;
                gw[p] = pack_bf16(__uint_as_float(dw << 16) * f0, __uint_as_float(dw & 0xFFFF0000u) * f1);
              }
#else
              gw[p] = mul_bf16x2(dAggp[h * (QH / 2) + p], sel_pair(prmt(zw, 0u, 0xbb99u), U_NEG, U_POS));
#endif
              if (DROP) gw[p] &= keep_pair(h ? kb.w : kb.z, i);
            }
            if (h == 0) sb.x = sw; else sb.y = sw;
            // in place over D2: both packed halves land in fp32 columns [0,96), which round h = 0 reads
            if (h == 0) named_bar_sync(bar_id, 128);
#pragma unroll
            for (int c = 0; c < QH / 8; ++c)
              tmem_st4(tp + C_RB + h * (NH2 / 2) + 16 * c, gw[4 * c], gw[4 * c + 1], gw[4 * c + 2], gw[4 * c + 3]);
            tmem_st_wait();
            tc_fence_before();
            warp_arrive(bar0 + (h ? BwdBars::rdyC1 : BwdBars::rdyC));   // this half of G2' feeds M3's K steps 6h..6h+5
          }
          t.sbits[(size_t)(g0 + it) * F_NEPI + threadIdx.x] = sb;
          MPG_TRW(it, 4);
        } else {
          const uint2 sb = t.sbits[(size_t)(g0 + it) * F_NEPI + threadIdx.x];
          uint32_t g2w[Q2 / 2];
#pragma unroll
          for (int h = 0; h < 2; ++h)
#pragma unroll
            for (int p = 0; p < QH / 2; ++p) {
              uint32_t gw = mul_bf16x2(dAggp[h * (QH / 2) + p], sel_pair(neg_pair_mask(h ? sb.y : sb.x, p), U_NEG, U_POS));
              if (DROP) gw &= keep_pair(h ? kb.w : kb.z, 2 * p);
              g2w[h * (QH / 2) + p] = gw;
            }
          if (it >= 1) mbar_wait(bar0 + BwdBars::doneC, (it - 1) & 1);   // M6(it-1) done: G2' tile free
          // chunk 12h + 4c + q of the G2' tile = chunk 4(c + 3h) + q
#pragma unroll
          for (int h = 0; h < 2; ++h)
#pragma unroll
            for (int c = 0; c < QH / 8; ++c) {
              const int cc = c + 3 * h;
              asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(((cc & 1) ? xs1 : xs0) + D_OFF_G2 + (cc >> 1) * A_BLK),
                           "r"(g2w[h * 12 + 4 * c]), "r"(g2w[h * 12 + 4 * c + 1]), "r"(g2w[h * 12 + 4 * c + 2]),
                           "r"(g2w[h * 12 + 4 * c + 3]));
            }
          fence_async_smem();
          warp_arrive(bar0 + BwdBars::rdyC);
        }

        if constexpr (CH) {
          // ---- E3: G1' = dH1 * keep1 * (1 + cg sgn D1) -> TMEM in place (A of M4) and shared memory (B of M5) ---------
          mbar_wait(bar0 + BwdBars::doneC, par);
          MPG_TRW(it, 5);
          tc_fence_after();
          auto e3_round = [&](auto RND) {
            constexpr int rnd = decltype(RND)::value, nc = rnd == 0 ? 3 : 2, c0 = rnd * 3;
            float v[8 * nc];
            if constexpr (rnd == 0) tmem_ld8x3(tl + C_RA, v);
            else tmem_ld8x2(tl + C_RA + 96, v);
            uint32_t w[4 * nc];
#pragma unroll
            for (int i = 0; i < 8 * nc; i += 2) {
              const int el = 8 * c0 + i, p = el >> 1;
#ifdef MPG_FP32_SLOPE
              {
                const uint32_t a_bits = __float_as_uint(a.alpha), one_bits = 0x3F800000u;
                const float f0 = __uint_as_float(sel_pair(neg_lane_mask(s1[p >> 4], p & 15, 0), a_bits, one_bits));
                const float f1 = __uint_as_float(sel_pair(neg_lane_mask(s1[p >> 4], p & 15, 1), a_bits, one_bits));
                w[i >> 1] = pack_bf16(v[i] * f0, v[i + 1] * f1);
              }
#else
              w[i >> 1] = mul_bf16x2(pack_bf16(v[i], v[i + 1]), sel_pair(neg_pair_mask(s1[p >> 4], p & 15), F_NEG, F_POS));
#endif
              if (DROP) w[i >> 1] &= keep_pair(el < 32 ? kb.x : kb.y, el & 31);
            }
            if constexpr (rnd == 0) named_bar_sync(bar_id, 128);
#pragma unroll
            for (int c = 0; c < nc; ++c) {
              const int cc = c0 + c;
              tmem_st4(tp + C_RA + 16 * cc, w[4 * c], w[4 * c + 1], w[4 * c + 2], w[4 * c + 3]);
              asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(((cc & 1) ? xs1 : xs0) + C_OFF_X + (cc >> 1) * A_BLK),
                           "r"(w[4 * c]), "r"(w[4 * c + 1]), "r"(w[4 * c + 2]), "r"(w[4 * c + 3]));
            }
          };
          e3_round(std::integral_constant<int, 0>{});
          e3_round(std::integral_constant<int, 1>{});
          tmem_st_wait();
          fence_async_smem();
          tc_fence_before();
          warp_arrive(bar0 + BwdBars::rdyD);
          MPG_TRW(it, 6);

          // ---- H0' of the next step under M4 (M5 has released the H0' tile and X) ---------------------------------------
          mbar_wait(bar0 + BwdBars::doneD5, par);
          MPG_TRW(it, 7);
          if (it + 1 < nsteps) build_h0(it + 1);
          MPG_TRW(it, 8);

          // ---- E4: G0' = dH0 * keep0 * (1 + cg sgn pre0): dP in registers, bf16 tile for the dQ MMA ----------------------
          mbar_wait(bar0 + BwdBars::doneD4, par);
          MPG_TRW(it, 9);
          tc_fence_after();
          {
            float v[Q0];
            tmem_ld8x3(tl + C_RB + NH2, v);
            const uint32_t a_bits = __float_as_uint(a.alpha), one_bits = 0x3F800000u;
#pragma unroll
            for (int c = 0; c < Q0 / 8; ++c) {
              uint32_t w[4];
#pragma unroll
              for (int e = 0; e < 8; e += 2) {
                const int p = 4 * c + (e >> 1);
                float g[2];
#pragma unroll
                for (int hi = 0; hi < 2; ++hi) {
                  uint32_t f = sel_pair(neg_lane_mask(s0, p, hi), a_bits, one_bits);
                  if (DROP) f &= keep_one(k0w, 8 * c + e + hi);
                  g[hi] = v[8 * c + e + hi] * __uint_as_float(f);
                  dPacc[8 * c + e + hi] += g[hi];
                }
                w[e >> 1] = pack_bf16(g[0], g[1]);
              }
              asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(((c & 1) ? xs1 : xs0) + C_OFF_X + (c >> 1) * A_BLK),
                           "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]));
            }
          }
          fence_async_smem();
          tc_fence_before();
          warp_arrive(bar0 + BwdBars::rdyE);
          MPG_TRW(it, 10);
          s0 = s0_next;
          k0w = k0w_next;
          kb = n_kb;                          // draw of step it+1 (build_h0(it + 1) ran above)
          // remember where this step's dQ sums belong, then advance
          p_j0 = c_j0; p_nj = c_nj;
          p_s = c_s;
          {   // advance; flush dP when the next step belongs to another tile (or there is none)
            const int2 nx = it + 1 < nsteps ? make_int2(n_tile, n_s) : make_int2(-1, 0);   // build_h0(it + 1) ran above
            c_first = nx.x != c_tile;
            c_s = nx.y;
            c_m = n_m;
            if (c_first) {
              if (c_valid) {
                // (layouts as in tile_rows: tiled dP makes each warp-wide reduction 512 contiguous bytes)
                float* dst = a.p_tiled ? a.dP + (size_t)c_tile * TILE * K0 + ((size_t)(2 * q) * TILE + row) * 4
                                       : a.dP + (size_t)c_r * K0 + q * 8;
                const int cs = a.p_tiled ? 8 * TILE * 4 : 32, hs = a.p_tiled ? TILE * 4 : 4;
#pragma unroll
                for (int c = 0; c < Q0; c += 4)
                  red_add_v4(dst + cs * (c >> 3) + hs * ((c >> 2) & 1), dPacc[c] * sc_g0, dPacc[c + 1] * sc_g0,
                             dPacc[c + 2] * sc_g0, dPacc[c + 3] * sc_g0);
              }
#pragma unroll
              for (int c = 0; c < Q0; ++c) dPacc[c] = 0.f;
              c_tile = nx.x;
              if (it + 1 < nsteps) enter_tile();
            }
          }
        } else {
          // ---- DW2: H0' two steps ahead (layer 1 runs one step ahead of the dW2 MMA) ---------------------------------------
          if (it + 2 < nsteps) {
            mbar_wait(bar0 + BwdBars::doneA, (it + 1) & 1);   // M1(it+1) done: H0' tile free
            build_h0(it + 2);
          }
          if (it + 1 < nsteps) {
            c_s = p1_s;
            c_m = p1_m;
            if (p1_tile != c_tile) {
              c_tile = p1_tile;
              enter_tile();
            }
            kb = p1_kb;
            p1_tile = n_tile; p1_s = n_s; p1_m = n_m; p1_kb = n_kb;
          }
        }
      }

      // ---- drain -----------------------------------------------------------------------------------------------------------
      MPG_TP(3);
      if constexpr (CH) {
        mbar_wait(bar0 + BwdBars::doneE, (nsteps - 1) & 1);
        tc_fence_after();
        MPG_TP(4);
        dq_readout();
        MPG_TP(5);
        // dW1^T accumulator: lane = H0' column k0 (< 96: dW1[:, k0]; 96: db1), column n1 = 32c + 8q + e
        float* slab = t.wslab + (size_t)blockIdx.x * SLAB_FLOATS;
        float v[Q1];
        tmem_ld8x5(tl + C_PW1, v);
#pragma unroll
        for (int i = 0; i < Q1; ++i) {
          const int n1 = 32 * (i >> 3) + 8 * q + (i & 7);
          if (row < K0) slab[SLAB_DW1 + n1 * K0 + row] = v[i] * sc_dw1;
          else if (row == K0) slab[SLAB_DB1 + n1] = v[i] * sc_db1;
        }
      } else {
        mbar_wait(bar0 + BwdBars::doneC, (nsteps - 1) & 1);
        tc_fence_after();
        float* slab = t.wslab + (size_t)blockIdx.x * SLAB_FLOATS;
#pragma unroll 1
        for (int mb = 0; mb < 2; ++mb) {
          const int n2 = mb * 128 + row;
#pragma unroll 1
          for (int c = 0; c < 6; ++c) {
            const int col = 32 * c + 8 * q;   // accumulator column n1 (160 = db2)
            if (col >= NDW) continue;         // warp-uniform
            float v[8];
            tmem_ld8(tmem + ((uint32_t)((warp & 3) * 32) << 16) + (mb ? D_PW2B : D_PW2A) + col, v);
            if (n2 < N2) {
#pragma unroll
              for (int e = 0; e < 8; ++e) {
                if (col + e < N1) slab[SLAB_DW2 + (col + e) * N2 + n2] = v[e] * sc_dw2;   // [n1][n2]: lanes = n2, coalesced
                else if (col + e == N1) slab[SLAB_DB2 + n2] = v[e] * sc_db2;
              }
            }
          }
        }
      }
    }
  }

  MPG_TP(6);
  tc_fence_before();
  __syncthreads();
  MPG_TP(7);
  if (warp == 16) tmem_dealloc(tmem, TMEM_COLS);
}

// dW1 / db1 / dW2 / db2 += sum over the CTAs' slabs.  27 MB of slabs at N = 30: the loop over slabs must keep many
// loads in flight, so a block is 32 float4 columns x 8 slab groups (slab c goes to group c % 8), ~19 independent
// 16-byte loads per thread; the groups are combined in a fixed order (deterministic).
constexpr int WR_COLS = 32, WR_PARTS = 8;
__global__ void __launch_bounds__(WR_COLS * WR_PARTS) wgrad_reduce_kernel(const float* __restrict__ slabs, int nslabs,
                                                                          float* __restrict__ dW1, float* __restrict__ db1,
                                                                          float* __restrict__ dW2, float* __restrict__ db2) {
  static_assert(SLAB_FLOATS % 4 == 0, "slabs are read as float4");
  __shared__ float4 part[WR_PARTS][WR_COLS];
  const int lane = threadIdx.x & (WR_COLS - 1), p = threadIdx.x / WR_COLS;
  const int c4 = blockIdx.x * WR_COLS + lane;
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
  if (c4 < SLAB_FLOATS / 4) {
#pragma unroll 4
    for (int c = p; c < nslabs; c += WR_PARTS) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(slabs + (size_t)c * SLAB_FLOATS) + c4);
      s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
    }
  }
  part[p][lane] = s;
  __syncthreads();
  if (p != 0 || c4 >= SLAB_FLOATS / 4) return;
#pragma unroll
  for (int k = 1; k < WR_PARTS; ++k) {
    const float4 v = part[k][lane];
    s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
  }
  const float sv[4] = {s.x, s.y, s.z, s.w};
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const int i = 4 * c4 + e;
    float* dst;
    if (i < SLAB_DB1) dst = dW1 + i;
    else if (i < SLAB_DW2) dst = db1 + (i - SLAB_DB1);
    else if (i < SLAB_DB2) dst = dW2 + ((i - SLAB_DW2) % N2) * N1 + (i - SLAB_DW2) / N2;   // slab holds dW2 transposed
    else dst = db2 + (i - SLAB_DB2);
    *dst += sv[e];
  }
}
