// One-shot gradient all-reduce fused with the RMSprop update over NVLink peer memory: ONE kernel
// per network and step replaces ncclAllReduce + the optimizer kernel.
//
// Every rank's flat gradient buffer lives in symmetric memory (torch.distributed._symmetric_memory: the same
// allocation mapped into every peer); `grads[r]` is rank r's buffer as seen from this GPU, `flags[r]` rank r's flag
// block.  CTA c of every rank owns the same contiguous chunk of the flat parameter vector:
//
//   1. arrive:  flags[p][A, c, me] += 1 on every peer p (release, system scope), then wait until my own
//      flags[me][A, c, p] reach this call's epoch for every p (acquire): rank p's kernel is running, so the backward
//      kernels that produced its gradients -- earlier in its stream -- have completed.
//   2. (grads_mc != null: g = multimem.ld_reduce over the multicast mapping -- the NVSwitch sums the ranks' buffers,
//      each rank reads 1.4 MB instead of W x 1.4 MB; else:)  g = sum_r grads[r][i] in RANK ORDER (so every rank computes bit-identical sums, hence bit-identical weights),
//      read from the peers with 16-byte system-scope loads; sq = a sq + (1-a) (g/W)^2; p -= lr (g/W) / (sqrt(sq) + eps).
//      The 1.4 MB buffers make the one-shot form (every rank reads everything: (W-1) x 1.4 MB over NVLink) cheaper
//      than a reduce-scatter + all-gather pair.
//   3. depart:  flags[p][B, c, me] += 1, wait for flags[me][B, c, p]: every peer has finished reading my chunk c, so
//      the next kernel in my stream may overwrite the gradient buffer.
//
// The epoch lives on the device (`epoch[0]`, bumped by the last CTA to finish) so a captured CUDA graph replays it.
#include "common.cuh"
#include "peer.cuh"

namespace mpg {
namespace {

__device__ __forceinline__ void red_release_sys(unsigned* addr, unsigned v) {
  asm volatile("red.release.sys.global.add.u32 [%0], %1;" ::"l"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* addr) {
  unsigned v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(addr) : "memory");
  return v;
}
__device__ __forceinline__ float4 ld_sys_v4(const float* addr) {
  float4 v;
  asm volatile("ld.relaxed.sys.global.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(addr) : "memory");
  return v;
}
// sum over all ranks of the 16 bytes at this multicast address, reduced inside the NVSwitch (NVLink SHARP)
__device__ __forceinline__ float4 ld_reduce_mc_v4(const float* mc_addr) {
  float4 v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(mc_addr) : "memory");
  return v;
}
__device__ __forceinline__ float ld_sys(const float* addr) {
  float v;
  asm volatile("ld.relaxed.sys.global.f32 %0, [%1];" : "=f"(v) : "l"(addr) : "memory");
  return v;
}

// flag block of one rank: [phase (2)][cta][sender rank] unsigned counters, then the epoch word and the finished-CTA
// counter (local use only)
__device__ __forceinline__ size_t flag_index(int phase, int cta, int sender, int ctas, int world) {
  return ((size_t)phase * ctas + cta) * world + sender;
}

__device__ __forceinline__ void peer_barrier(const PeerArgs& a, int phase, unsigned epoch) {
  // one warp does the exchange: lane r talks to rank r
  if (threadIdx.x < 32) {
    const int r = threadIdx.x;
    if (r < a.world && r != a.rank)
      red_release_sys(a.flags[r] + flag_index(phase, blockIdx.x, a.rank, gridDim.x, a.world), 1u);
    if (r < a.world && r != a.rank) {
      const unsigned* mine = a.flags[a.rank] + flag_index(phase, blockIdx.x, r, gridDim.x, a.world);
      unsigned spins = 0;
      while ((int)(ld_acquire_sys(mine) - epoch) < 0) {
        if (++spins > (1u << 26)) __trap();   // a protocol bug must trap, never hang the GPU
      }
    }
  }
  __syncthreads();
}

constexpr int PEER_THREADS = 512, PEER_UNROLL = 4;

// 512 threads x 4 independent 16-byte loads per peer and pass keep the NVLink latency covered.  Measured on 2 B200s
// for the 1.42 MB buffer (profiles/r2_bench_peer_2gpu.txt): 123 / 68 / 40 / 29 / 24 / 21 us with 2 / 4 / 8 / 16 / 32 /
// 64 CTAs against 32 us for ncclAllReduce + the RMSprop kernel (9 us of it): the trainer launches 64.  A waiting CTA
// spins on its flags; its registers keep a persistent edge-kernel CTA of the other stream (~61 k registers) off that
// SM meanwhile, which is harmless here because D's update overlaps only the small node-level kernels that open
// train_G.
__global__ void __launch_bounds__(PEER_THREADS) allreduce_rmsprop_kernel(PeerArgs a) {
  unsigned* ctl = a.flags[a.rank] + 2 * (size_t)gridDim.x * a.world;   // [0] epoch, [1] finished CTAs
  const unsigned epoch = ld_acquire_sys(ctl) + 1u;   // read by every CTA before any CTA can bump it (see below)
  peer_barrier(a, 0, epoch);

  const size_t per = (((a.n + gridDim.x - 1) / gridDim.x) + 3) & ~(size_t)3;
  const size_t lo = (size_t)blockIdx.x * per, hi = min(a.n, lo + per);
  const float gs = 1.f / (float)a.world;
  constexpr size_t PASS = (size_t)PEER_THREADS * PEER_UNROLL * 4;
  for (size_t base = lo; base < hi; base += PASS) {
    float4 g[PEER_UNROLL];
    size_t idx[PEER_UNROLL];
#pragma unroll
    for (int u = 0; u < PEER_UNROLL; ++u) {
      idx[u] = base + ((size_t)u * PEER_THREADS + threadIdx.x) * 4;
      g[u] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    if (a.grads_mc != nullptr) {   // in-switch reduction: every rank reads the buffer ONCE instead of world times
#pragma unroll
      for (int u = 0; u < PEER_UNROLL; ++u)
        if (idx[u] + 4 <= hi) g[u] = ld_reduce_mc_v4(a.grads_mc + idx[u]);
    } else {
      for (int r = 0; r < a.world; ++r) {   // rank order: bit-identical sums on every rank
        float4 v[PEER_UNROLL];
#pragma unroll
        for (int u = 0; u < PEER_UNROLL; ++u)
          v[u] = idx[u] + 4 <= hi ? ld_sys_v4(a.grads[r] + idx[u]) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int u = 0; u < PEER_UNROLL; ++u) { g[u].x += v[u].x; g[u].y += v[u].y; g[u].z += v[u].z; g[u].w += v[u].w; }
      }
    }
#pragma unroll
    for (int u = 0; u < PEER_UNROLL; ++u) {
      const size_t i = idx[u];
      if (i + 4 <= hi) {
        float4 p = *reinterpret_cast<float4*>(a.p + i), s = *reinterpret_cast<float4*>(a.sq + i);
        const float gv[4] = {g[u].x * gs, g[u].y * gs, g[u].z * gs, g[u].w * gs};
        float* pp = &p.x;
        float* sp = &s.x;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          sp[e] = a.alpha * sp[e] + (1.f - a.alpha) * gv[e] * gv[e];
          pp[e] -= a.lr * gv[e] / (sqrtf(sp[e]) + a.eps);
        }
        *reinterpret_cast<float4*>(a.p + i) = p;
        *reinterpret_cast<float4*>(a.sq + i) = s;
      } else {
        for (size_t j = i; j < hi; ++j) {   // ragged tail of the chunk (at most 3 elements)
          float gj = 0.f;
          for (int r = 0; r < a.world; ++r) gj += ld_sys(a.grads[r] + j);
          gj *= gs;
          const float sj = a.alpha * a.sq[j] + (1.f - a.alpha) * gj * gj;
          a.sq[j] = sj;
          a.p[j] -= a.lr * gj / (sqrtf(sj) + a.eps);
        }
      }
    }
  }
  __syncthreads();
  peer_barrier(a, 1, epoch);
  // The last CTA to get here publishes the new epoch for the next call.  Every CTA read the old value at its start,
  // and the bump needs ALL CTAs to have finished, so no CTA of this launch can see the new one.
  if (threadIdx.x == 0) {
    __threadfence();
    if (atomicAdd(ctl + 1, 1u) == gridDim.x - 1) {
      ctl[1] = 0u;
      __threadfence();
      atomicExch(ctl, epoch);
    }
  }
}

}  // namespace

size_t peer_flag_words(int ctas, int world) { return 2 * (size_t)ctas * world + 2; }

int launch_allreduce_rmsprop(const PeerArgs& a, int ctas, cudaStream_t s) {
  MPG_CHECK(a.world >= 2 && a.world <= MPG_PEER_MAX && a.rank >= 0 && a.rank < a.world, "allreduce_rmsprop: bad rank/world");
  MPG_CHECK(ctas > 0 && ctas <= 148, "allreduce_rmsprop: CTA count must be in [1, 148] (all CTAs must be co-resident)");
  for (int r = 0; r < a.world; ++r)
    MPG_CHECK(a.grads[r] != nullptr && a.flags[r] != nullptr && (reinterpret_cast<uintptr_t>(a.grads[r]) & 15) == 0,
              "allreduce_rmsprop: null / unaligned peer pointer");
  if (a.n == 0) return 0;
  allreduce_rmsprop_kernel<<<ctas, PEER_THREADS, 0, s>>>(a);
  MPG_LAUNCH_CHECK();
  return 0;
}

}  // namespace mpg
