// Forward node-level end of the factorised first edge layer as a device function (one 128-row tile per block): used by
// the stand-alone pq_fwd_kernel (edge_node.cu) and by the tcgen05 path's set-up kernel (edge_tc.cu), which runs it
// next to the weight-image and work-list blocks of the same launch.
//   P = x Wa^T + b0,  Q = x Wb^T      (W0 = [Wa | Wb | Wef], mpgan/model.py:77-83, 294-311)
#pragma once
#include "edge.cuh"

namespace mpg {

struct PqFwdArgs {
  const float* x; int ldx;
  const float* W0; int ldw;
  const float* b0;
  float* P; float* Q;
  int BN, F, H0, p_tiled;
  const int* cmap; int ctiles_max;   // receiver compaction (EdgeArgs::cmap): P tiles follow the map, Q stays per padded row
};
constexpr int PQ_ROWS = 128, PQ_NT = 256;

// dst(i) = src(i) for i < n over the block's NTHR threads, U loads issued before the first store
template <int U, int NTHR, class LoadF, class StoreF>
__device__ __forceinline__ void batched_fill(int n, LoadF&& ld, StoreF&& st) {
  for (int i0 = threadIdx.x; i0 < n; i0 += NTHR * U) {
    float v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int idx = i0 + u * NTHR;
      v[u] = idx < n ? ld(idx) : 0.f;
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int idx = i0 + u * NTHR;
      if (idx < n) st(idx, v[u]);
    }
  }
}

__host__ __device__ inline size_t pq_fwd_smem(int F, int H0) {
  return (size_t)(F * (H0 + 4) + PQ_ROWS * (F + 1) + PQ_ROWS * (H0 + 4)) * sizeof(float);
}

// One block = one 128-row tile x one output (job = 2 * tile + h: h = 0 -> P, 1 -> Q), 256 threads.  Thread = two rows
// (r, r + 64) x every fourth 8-column group: per input feature two x values and eight weights (two 16-byte
// broadcast loads) feed 16 FMAs.  Tiled P goes straight to global memory (a warp's 32 rows x 16 bytes are contiguous);
// row-major outputs go through a padded staging tile and leave coalesced.
__device__ __forceinline__ void pq_fwd_tile(const PqFwdArgs& a, int job, float* sm) {
  const int F = a.F, H0 = a.H0, BN = a.BN;
  const int tile = job >> 1, h = job & 1;
  const int H0P = H0 + 4, FP = F + 1;
  float* Ws = sm;                        // [F][H0P]    Ws[f][k] = W0[k][h*F + f]
  float* xs = Ws + F * H0P;              // [128][FP]
  float* st = xs + PQ_ROWS * FP;         // [128][H0P]  staging of a row-major output
  const int r0 = tile * PQ_ROWS;
  const bool mapped = a.cmap != nullptr && h == 0;      // P of a compacted tile: lane r stands for padded row map[r]
  if (mapped ? tile >= a.cmap[0] : r0 >= BN) return;    // (block-uniform)
  const int* map = mapped ? a.cmap + 2 + 2 * a.ctiles_max + tile * PQ_ROWS : nullptr;
  // (eight independent loads in flight per thread: a load -> store loop exposes one memory latency per iteration)
  batched_fill<8, PQ_NT>(F * H0,
      [&](int idx) { const int k = idx / F, f = idx % F; return a.W0[(size_t)k * a.ldw + h * F + f]; },
      [&](int idx, float v) { const int k = idx / F, f = idx % F; Ws[f * H0P + k] = v; });
  batched_fill<8, PQ_NT>(PQ_ROWS * F,
      [&](int idx) {
        const int r = idx / F, f = idx % F;
        const int gr = mapped ? map[r] : (r0 + r < BN ? r0 + r : -1);
        return gr >= 0 ? a.x[(size_t)gr * a.ldx + f] : 0.f;
      },
      [&](int idx, float v) { const int r = idx / F, f = idx % F; xs[r * FP + f] = v; });
  __syncthreads();
  const int ra = threadIdx.x & 63, rb = ra + 64, qt = threadIdx.x >> 6;
  const bool direct = h == 0 && a.p_tiled;
  for (int gp = qt; gp < H0 / 8; gp += 4) {             // 8-column groups of this output
    const int k0 = 8 * gp;
    float acc[2][8];
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[0][e] = acc[1][e] = h ? 0.f : a.b0[k0 + e];
    const float* w = Ws + k0;
    for (int f = 0; f < F; ++f) {
      const float xa = xs[ra * FP + f], xb = xs[rb * FP + f];
      const float4 w0 = *reinterpret_cast<const float4*>(w + f * H0P);
      const float4 w1 = *reinterpret_cast<const float4*>(w + f * H0P + 4);
      const float wv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        acc[0][e] = fmaf(xa, wv[e], acc[0][e]);
        acc[1][e] = fmaf(xb, wv[e], acc[1][e]);
      }
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int rl = i ? rb : ra;
      if (direct) {
        if (mapped || r0 + rl < BN) {   // (mapped: every lane of the tile gets a finite P, rows or not)
          float* p = a.P + p_tiled_index((size_t)(r0 + rl), k0, H0);
          *reinterpret_cast<float4*>(p) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
          *reinterpret_cast<float4*>(p + 4 * PQ_ROWS) = make_float4(acc[i][4], acc[i][5], acc[i][6], acc[i][7]);
        }
      } else {
        float* sp = st + rl * H0P + k0;
        *reinterpret_cast<float4*>(sp) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
        *reinterpret_cast<float4*>(sp + 4) = make_float4(acc[i][4], acc[i][5], acc[i][6], acc[i][7]);
      }
    }
  }
  if (!direct) {
    __syncthreads();
    const int c4n = H0 / 4;
    float* dst = h ? a.Q : a.P;
    for (int idx = threadIdx.x; idx < PQ_ROWS * c4n; idx += PQ_NT) {
      const int r = idx / c4n, c4 = idx % c4n;
      if (r0 + r < BN)
        *reinterpret_cast<float4*>(dst + (size_t)(r0 + r) * H0 + 4 * c4) =
            *reinterpret_cast<const float4*>(st + r * H0P + 4 * c4);
    }
  }
}

}  // namespace mpg
