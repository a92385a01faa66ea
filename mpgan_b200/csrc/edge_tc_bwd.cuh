// tcgen05 backward of the edge network (included by edge_tc.cu inside its anonymous namespace).
//
// The backward recomputes H0/H1/D2 per (tile, sender) step exactly as the forward does and never
// stores an N^2 x hidden tensor.  All fp32 weight-gradient accumulators cannot be TMEM-resident at
// once (dW2: 2x160 columns, dW1^T: 160 columns, plus >= 192 streaming columns > 512), so the work is
// split into two kernels that share the recompute code:
//
//   CHAIN  G2 = dAgg*m*f'(D2) -> dH1 = G2 W2 -> G1 = dH1 f'(D1) -> dH0 = G1 W1 -> G0 = dH0 f'(pre0)
//          dP[r] += G0 (registers), dQ[jet,s] += G0 (red.global), dW1^T += H0^T G1 (TMEM accumulator;
//          the constant-1 bias columns of the H0 tile make row 96 of it db1).
//   DW2    recompute -> G2, dW2 += G2^T H1 (two M=128 TMEM accumulators), db2 = column sums of G2.
//
// The transposed products reuse the SAME swizzled shared-memory tiles through MN-major UMMA
// descriptors (a K-major [rows x cols] SW128 tile read as MN-major is its transpose), and the
// weight images serve both W (K-major B operand) and W^T (MN-major B operand).
//
// Scale convention with dropout p = 0.5 (s = 2): tiles hold activations / pre-activation gradients
// divided by s and the weight images hold s*W, so every MMA sees the true product; the factors are
// restored at the flush (dW: s^2, db: s) and in G0 (s).
//
// Thread layout as in the forward kernel: 16 warps, thread <-> (tile row, column quarter).

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

constexpr uint32_t BW_OFF_A0_CHAIN = OFF_W2 + W2_BYTES;            // 114688: H0 tile (32 KB)
constexpr uint32_t BW_OFF_A1_CHAIN = BW_OFF_A0_CHAIN + H0_BYTES;   // 147456: H1 / G2 / G1 tile (48 KB)
constexpr uint32_t BW_OFF_G2_DW2 = OFF_W2 + W2_BYTES;              // 114688: H0 then G2 tile (48 KB)
constexpr uint32_t BW_OFF_A1_DW2 = BW_OFF_G2_DW2 + H1_BYTES;       // 163840: H1 tile (48 KB)
constexpr uint32_t BW_OFF_BAR_CHAIN = BW_OFF_A1_CHAIN + H1_BYTES;  // 196608
constexpr uint32_t BW_OFF_BAR_DW2 = BW_OFF_A1_DW2 + H1_BYTES;      // 212992
constexpr uint32_t BW_SMEM_CHAIN = BW_OFF_BAR_CHAIN + 128 + 1024;
constexpr uint32_t BW_SMEM_DW2 = BW_OFF_BAR_DW2 + 128 + 1024;
constexpr uint32_t RS_COL = 0, RW1_COL = 256, RW2A_COL = 192, RW2B_COL = 352;

template <int MODE, bool DROP>
__global__ void __launch_bounds__(NTHREADS, 1) edge_tc_bwd_kernel(TcArgs t) {
  extern __shared__ uint8_t smem_raw[];
  const EdgeArgs& a = t.a;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* sm = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t sW1 = base + OFF_W1, sW2 = base + OFF_W2;
  const uint32_t sA0 = base + (MODE == BWD_CHAIN ? BW_OFF_A0_CHAIN : BW_OFF_G2_DW2);
  const uint32_t sA1 = base + (MODE == BWD_CHAIN ? BW_OFF_A1_CHAIN : BW_OFF_A1_DW2);
  const uint32_t sG2 = MODE == BWD_CHAIN ? sA1 : base + BW_OFF_G2_DW2;   // G2 tile (CHAIN: over H1; DW2: over H0)
  const uint32_t off_bar = MODE == BWD_CHAIN ? BW_OFF_BAR_CHAIN : BW_OFF_BAR_DW2;
  const uint32_t bar_w = base + off_bar, bar_rdy = bar_w + 8, bar_done = bar_w + 16;
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(sm + off_bar + 64);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  DropCfg drop = a.drop;
  if (DROP) resolve_seed(drop);
  const float sdrop = DROP ? 2.f : 1.f;

  const long long g0 = t.total_steps * blockIdx.x / gridDim.x;
  const long long g1 = t.total_steps * (blockIdx.x + 1) / gridDim.x;
  const int nsteps = (int)(g1 - g0);

  if (threadIdx.x == 0) {
    mbar_init(bar_w, 1);
    mbar_init(bar_rdy, NTHREADS);
    mbar_init(bar_done, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) tmem_alloc(smem_u32(const_cast<uint32_t*>(tmem_slot)), TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const bool issuer = threadIdx.x == 0;
  if (issuer && nsteps > 0) {
    mbar_expect_tx(bar_w, W1_BYTES + W2_BYTES);
    bulk_g2s(sW1, t.w1img, W1_BYTES, bar_w);
    bulk_g2s(sW2, t.w2img, W2_BYTES, bar_w);
  }

  const int q = warp >> 2;
  const int row = (warp & 3) * 32 + lane;
  const uint32_t tlane = (uint32_t)((warp & 3) * 32) << 16;
  const int BN = a.B * a.N;
  uint32_t prdy = 0, pdone = 0;

  // publish this thread's smem writes + TMEM reads, let thread 0 issue a batch of MMAs, wait for them
  auto sync_issue = [&](auto&& issue_fn) {
    fence_async_smem();
    tc_fence_before();
    mbar_arrive(bar_rdy);
    if (issuer) {
      mbar_wait(bar_rdy, prdy);
      tc_fence_after();
      issue_fn();
      umma_commit(bar_done);
    }
    prdy ^= 1;
    mbar_wait(bar_done, pdone);
    pdone ^= 1;
    tc_fence_after();
  };

  // CHAIN: the H0 tile is private to H0; its bias columns (96,97 = 1) and the never-written columns
  // 98..127 (read as unused rows by the transposed dW1 product) are initialised once
  if (MODE == BWD_CHAIN && nsteps > 0 && q == 0) {
    st_ones_chunk(sA0 + swz_chunk(row, 96, A_BLK));
    st_zero_chunk(sA0 + swz_chunk(row, 104, A_BLK));
    st_zero_chunk(sA0 + swz_chunk(row, 112, A_BLK));
    st_zero_chunk(sA0 + swz_chunk(row, 120, A_BLK));
  }

  float Preg[Q0];
  uint32_t dAggp[Q2 / 2];       // bf16x2 pairs of dAgg[r][h*96 + q*24 + ..], h = 0, 1
  float dPacc[MODE == BWD_CHAIN ? Q0 : 1];
  float db2acc0 = 0.f, db2acc1 = 0.f;
  int cur_tile = -1, r = 0, jet = 0;
  bool valid = false;
  bool first_mma = true;

  auto flush_dP = [&]() {
    if constexpr (MODE == BWD_CHAIN) {
      if (cur_tile >= 0 && valid) {
        float* dst = a.dP + (size_t)r * K0 + q * 8;
#pragma unroll
        for (int c = 0; c < Q0; ++c) atomicAdd(dst + 32 * (c >> 3) + (c & 7), dPacc[c]);
      }
    }
  };
  auto load_tile = [&](int tile) {
    flush_dP();
    cur_tile = tile;
    r = tile * TILE + row;
    valid = r < BN;
    const int rc = valid ? r : BN - 1;
    jet = rc / a.N;
    // column ownership: chunks 4c + q, i.e. columns 32c + 8q + [0, 8) (edge_tc_common.cuh)
#pragma unroll
    for (int c = 0; c < Q0 / 4; ++c) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(a.P + (size_t)rc * K0 + q * 8 + 32 * (c >> 1) + 4 * (c & 1)));
      Preg[4 * c] = v.x; Preg[4 * c + 1] = v.y; Preg[4 * c + 2] = v.z; Preg[4 * c + 3] = v.w;
    }
#pragma unroll
    for (int h = 0; h < 2; ++h) {   // this thread's layer-2 columns: h*96 + 32c + 8q + [0, 8)
      const float* dg = a.dagg + (size_t)rc * N2 + h * NH2 + q * 8;
#pragma unroll
      for (int c = 0; c < QH / 4; ++c) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(dg + 32 * (c >> 1) + 4 * (c & 1)));
        dAggp[h * (QH / 2) + 2 * c] = pack_bf16(v.x, v.y);
        dAggp[h * (QH / 2) + 2 * c + 1] = pack_bf16(v.z, v.w);
      }
    }
    if constexpr (MODE == BWD_CHAIN) {
#pragma unroll
      for (int c = 0; c < Q0; ++c) dPacc[c] = 0.f;
    }
  };

  if (issuer && nsteps > 0) mbar_wait(bar_w, 0);

  for (int it = 0; it < nsteps; ++it) {
    const long long g = g0 + it;
    const int tile = (int)(g / a.N), s = (int)(g % a.N);
    if (tile != cur_tile) load_tile(tile);
    const uint64_t pair = (uint64_t)(valid ? r : 0) * a.N + s;
    const float mfac = valid ? (a.mask ? a.mask[(size_t)jet * a.N + s] : 1.f) * a.out_scale : 0.f;
    uint32_t k0w = 0;           // keep words (common.cuh edge_drop_*): layer 0
    u4 bits{0, 0, 0, 0};        // x,y: layer 1; z / w: layer 2 low / high half
    if (DROP) {
      k0w = edge_drop_bits(drop.seed, pair, q, 0).x;
      bits = edge_drop_bits(drop.seed, pair, q, 1);
    }

    // ---- H0 tile; remember the sign bits of the 24 columns this thread owns -----------------------------
    uint32_t pos0 = 0;
    {
      const float4* qp = reinterpret_cast<const float4*>(a.Q + ((size_t)jet * a.N + s) * K0 + q * 8);
#pragma unroll
      for (int c8 = 0; c8 < Q0 / 8; ++c8) {
        const float4 q0 = __ldg(qp + 8 * c8), q1 = __ldg(qp + 8 * c8 + 1);
        float v[8] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w};
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const int le = c8 * 8 + e;
          float x = v[e] + Preg[le];
          if (x > 0.f) pos0 |= 1u << le;
          x = fmaxf(x, a.alpha * x);
          if (DROP) x = apply_keep(x, keep_one(k0w, le));
          v[e] = x;
        }
        st_chunk(sA0 + swz_chunk(row, 32 * c8 + 8 * q, A_BLK), v);
      }
      if (MODE == BWD_DW2 && q == 0) {   // DW2 builds H0 inside the G2 region: bias columns every step
        st_ones_chunk(sA0 + swz_chunk(row, 96, A_BLK));
        st_zero_chunk(sA0 + swz_chunk(row, 104, A_BLK));
      }
    }
    sync_issue([&]() {   // D1 = H0 * W1^T
#pragma unroll
      for (int ks = 0; ks < KSTEPS1; ++ks) {
        const uint32_t blk = ks >> 2, j = ks & 3;
        umma_bf16(tmem + RS_COL, umma_desc(sA0 + blk * A_BLK + j * 32), umma_desc(sW1 + blk * W1_BLK + j * 32),
                  umma_idesc(N1), ks > 0);
      }
    });
    // ---- e1: H1 tile, sign bits of this thread's 40 columns ----------------------------------------------
    uint32_t pos1[2] = {0, 0};
    {
      float v[Q1];
      tmem_ld8x5(tmem + tlane + RS_COL + q * 8, v);
#pragma unroll
      for (int le = 0; le < Q1; ++le) {
        float x = v[le];
        if (x > 0.f) pos1[le >> 5] |= 1u << (le & 31);
        x = fmaxf(x, a.alpha * x);
        if (DROP) x = apply_keep(x, keep_one(le < 32 ? bits.x : bits.y, le & 31));
        v[le] = x;
      }
#pragma unroll
      for (int c = 0; c < Q1 / 8; ++c) st_chunk(sA1 + swz_chunk(row, 32 * c + 8 * q, A_BLK), v + 8 * c);
      // constant bias columns of the H1 tile: DW2 never overwrites them, CHAIN reuses the region for G2/G1
      if ((MODE == BWD_CHAIN || it == 0) && q == 1) {
        st_ones_chunk(sA1 + swz_chunk(row, 160, A_BLK));
        st_zero_chunk(sA1 + swz_chunk(row, 168, A_BLK));
      }
    }
    sync_issue([&]() {   // D2 = H1 * W2^T
#pragma unroll
      for (int ks = 0; ks < KSTEPS2; ++ks) {
        const uint32_t blk = ks >> 2, j = ks & 3;
        umma_bf16(tmem + RS_COL, umma_desc(sA1 + blk * A_BLK + j * 32), umma_desc(sW2 + blk * W2_BLK + j * 32),
                  umma_idesc(N2), ks > 0);
      }
    });
    // ---- e2': G2 = dAgg * m * f'(D2) * keep2  (divided by s, see header) -> bf16 tile --------------------
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      float v[QH];
      tmem_ld8x3(tmem + tlane + RS_COL + h * NH2 + q * 8, v);
#pragma unroll
      for (int e = 0; e < QH; ++e) {
        const uint32_t pk = dAggp[h * (QH / 2) + (e >> 1)];
        const float dg = __uint_as_float((e & 1) ? (pk & 0xFFFF0000u) : (pk << 16));
        float gval = dg * mfac * (v[e] > 0.f ? 1.f : a.alpha);
        if (DROP) gval = apply_keep(gval, keep_one(h ? bits.w : bits.z, e));
        v[e] = gval;
      }
#pragma unroll
      for (int c8 = 0; c8 < QH / 8; ++c8) st_chunk(sG2 + swz_chunk(row, h * NH2 + 32 * c8 + 8 * q, A_BLK), v + c8 * 8);
    }

    if constexpr (MODE == BWD_DW2) {
      // db2 partial column sums straight from the bf16 G2 tile (thread -> 2 columns x 32 rows)
      asm volatile("bar.sync 1, %0;" ::"n"(NTHREADS) : "memory");
      if (threadIdx.x < 384) {
        const int cp = threadIdx.x % 96, rq = threadIdx.x / 96;
        const uint32_t col = 2 * cp;
#pragma unroll 8
        for (int rr = 0; rr < 32; ++rr) {
          const uint32_t rw = rq * 32 + rr;
          uint32_t w;
          asm volatile("ld.shared.b32 %0, [%1];" : "=r"(w) : "r"(sG2 + swz_chunk(rw, col, A_BLK) + (col & 7) * 2));
          db2acc0 += __uint_as_float(w << 16);
          db2acc1 += __uint_as_float(w & 0xFFFF0000u);
        }
      }
      const bool acc_flag = !first_mma;
      sync_issue([&]() {   // dW2[n2][n1] += G2^T H1 : two M=128 blocks over G2 columns, N = 160, K = 128 rows
#pragma unroll
        for (int mb = 0; mb < 2; ++mb)
#pragma unroll
          for (int ks = 0; ks < TILE / 16; ++ks)
            umma_bf16(tmem + (mb == 0 ? RW2A_COL : RW2B_COL), umma_desc_mn(sG2 + mb * 2 * A_BLK + ks * 2048, A_BLK),
                      umma_desc_mn(sA1 + ks * 2048, A_BLK), umma_idesc_t(N1, 1, 1), (acc_flag || ks > 0) ? 1u : 0u);
      });
      first_mma = false;
    } else {
      sync_issue([&]() {   // dH1 = G2 * W2   (B = W2 image read MN-major: N = 160 inputs, K = 192 outputs)
#pragma unroll
        for (int ks = 0; ks < N2 / 16; ++ks) {
          const uint32_t blk = ks >> 2, j = ks & 3;
          umma_bf16(tmem + RS_COL, umma_desc(sG2 + blk * A_BLK + j * 32), umma_desc_mn(sW2 + ks * 2048, W2_BLK),
                    umma_idesc_t(N1, 0, 1), ks > 0);
        }
      });
      // ---- e3: G1 = dH1 * f'(D1) * keep1 (divided by s) -> bf16 tile over the G2 tile --------------------
      {
        float v[Q1];
        tmem_ld8x5(tmem + tlane + RS_COL + q * 8, v);
#pragma unroll
        for (int le = 0; le < Q1; ++le) {
          float gval = v[le] * (((pos1[le >> 5] >> (le & 31)) & 1u) ? 1.f : a.alpha);
          if (DROP) gval = apply_keep(gval, keep_one(le < 32 ? bits.x : bits.y, le & 31));
          v[le] = gval;
        }
#pragma unroll
        for (int c = 0; c < Q1 / 8; ++c) st_chunk(sA1 + swz_chunk(row, 32 * c + 8 * q, A_BLK), v + 8 * c);
      }
      const bool acc_flag = !first_mma;
      sync_issue([&]() {
        // dH0 = G1 * W1   (B = W1 image read MN-major: N = 96 inputs, K = 160 outputs)
#pragma unroll
        for (int ks = 0; ks < N1 / 16; ++ks) {
          const uint32_t blk = ks >> 2, j = ks & 3;
          umma_bf16(tmem + RS_COL, umma_desc(sA1 + blk * A_BLK + j * 32), umma_desc_mn(sW1 + ks * 2048, W1_BLK),
                    umma_idesc_t(K0, 0, 1), ks > 0);
        }
        // dW1^T[k0][n1] += H0^T G1   (M = 128 columns of the H0 tile incl. the constant-1 columns, K = 128 rows)
#pragma unroll
        for (int ks = 0; ks < TILE / 16; ++ks)
          umma_bf16(tmem + RW1_COL, umma_desc_mn(sA0 + ks * 2048, A_BLK), umma_desc_mn(sA1 + ks * 2048, A_BLK),
                    umma_idesc_t(N1, 1, 1), (acc_flag || ks > 0) ? 1u : 0u);
      });
      first_mma = false;
      // ---- e4: G0 = dH0 * s * f'(pre0) * keep0 -> dP (registers), dQ (red.global) ---------------------------
      {
        float* dq = a.dQ + ((size_t)jet * a.N + s) * K0 + q * 8;
        float v[Q0];
        tmem_ld8x3(tmem + tlane + RS_COL + q * 8, v);
#pragma unroll
        for (int e = 0; e < Q0; ++e) {
          float gval = v[e] * sdrop * (((pos0 >> e) & 1u) ? 1.f : a.alpha);
          if (DROP) gval = apply_keep(gval, keep_one(k0w, e));
          if (valid) {
            dPacc[e] += gval;
            if (gval != 0.f) atomicAdd(dq + 32 * (e >> 3) + (e & 7), gval);
          }
        }
      }
    }
  }
  flush_dP();

  // ---- flush the TMEM weight-gradient accumulators ----------------------------------------------------
  if (nsteps > 0) {
    tc_fence_after();
    if constexpr (MODE == BWD_CHAIN) {
      // accumulator row m' = H0-tile column (k0 < 96: dW1[:, k0]; 96: db1), column n1
      float v[Q1];
      tmem_ld_cols<Q1>(tmem + tlane + RW1_COL + q * Q1, v);
#pragma unroll
      for (int e = 0; e < Q1; ++e) {
        const int n1 = q * Q1 + e;
        if (row < K0) atomicAdd(a.dW1 + (size_t)n1 * K0 + row, v[e] * sdrop * sdrop);
        else if (row == K0) atomicAdd(a.db1 + n1, v[e] * sdrop);
      }
    } else {
#pragma unroll 1
      for (int mb = 0; mb < 2; ++mb) {
        const int n2 = mb * 128 + row;
        float v[Q1];
        tmem_ld_cols<Q1>(tmem + tlane + (mb == 0 ? RW2A_COL : RW2B_COL) + q * Q1, v);
        if (n2 < N2) {
#pragma unroll
          for (int e = 0; e < Q1; ++e) atomicAdd(a.dW2 + (size_t)n2 * N1 + q * Q1 + e, v[e] * sdrop * sdrop);
        }
      }
      if (threadIdx.x < 384) {
        const int cp = threadIdx.x % 96;
        atomicAdd(a.db2 + 2 * cp, db2acc0 * sdrop);
        atomicAdd(a.db2 + 2 * cp + 1, db2acc1 * sdrop);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, TMEM_COLS);
}
