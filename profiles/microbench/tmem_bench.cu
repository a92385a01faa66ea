// Microbenchmark (B200, sm_100a): TMEM->register load throughput and the cost of candidate
// epilogue instruction sequences of the edge kernel, for 4 / 8 / 16 warps of one CTA.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_bench tmem_bench.cu && ./tmem_bench
#include <cstdio>
#include <cstdint>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void ld16(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}
__device__ __forceinline__ void ld32(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
      "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
}
__device__ __forceinline__ void ldwait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 p = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&p);
}

// mode 0: x16 ld + wait each; 1: 3 x16 lds then one wait; 2: x32 + wait; 3: E2 v1 (fmax/fmul + fma);
// 4: E2 v2 (fma |v| + fma); 5: E1 v1 (fmul,fmax,cvt,sts); 6: E1 v2 (fma|v|, cvt, sts); 7: E1 bf16x2 lrelu
// 8: E2 v2 with all three loads issued before one wait
template <int MODE>
__global__ void __launch_bounds__(512, 1) bench(int iters, float alpha, float m, float* out, long long* clk) {
  extern __shared__ uint8_t smem[];
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tmem = slot;
  const uint32_t tl = tmem + (((uint32_t)(warp & 3) * 32) << 16);
  const int q = (warp >> 2) & 3;
  const int row = (warp & 3) * 32 + lane;
  float acc[48];
#pragma unroll
  for (int i = 0; i < 48; ++i) acc[i] = 0.f;
  const float c = (1.f - alpha) / (1.f + alpha);
  const uint32_t sbase = (smem_u32(smem) + 1023u) & ~1023u;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0) {
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        uint32_t r[16];
        ld16(tl + q * 48 + k * 16 + (it & 1) * 256, r);
        ldwait();
#pragma unroll
        for (int e = 0; e < 16; ++e) acc[e] = __uint_as_float(__float_as_uint(acc[e]) ^ r[e]);
      }
    } else if (MODE == 1) {
      uint32_t r[48];
#pragma unroll
      for (int k = 0; k < 3; ++k) ld16(tl + q * 48 + k * 16 + (it & 1) * 256, r + 16 * k);
      ldwait();
#pragma unroll
      for (int e = 0; e < 48; ++e) acc[e] = __uint_as_float(__float_as_uint(acc[e]) ^ r[e]);
    } else if (MODE == 2) {
      uint32_t r[32];
      ld32(tl + q * 32 + (it & 1) * 256, r);
      ldwait();
#pragma unroll
      for (int e = 0; e < 32; ++e) acc[e] = __uint_as_float(__float_as_uint(acc[e]) ^ r[e]);
    } else if (MODE == 3) {
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        uint32_t r[16];
        ld16(tl + q * 48 + k * 16 + (it & 1) * 256, r);
        ldwait();
#pragma unroll
        for (int e = 0; e < 16; ++e) {
          const float v = __uint_as_float(r[e]);
          acc[k * 16 + e] = fmaf(fmaxf(v, alpha * v), m, acc[k * 16 + e]);
        }
      }
    } else if (MODE == 4) {
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        uint32_t r[16];
        ld16(tl + q * 48 + k * 16 + (it & 1) * 256, r);
        ldwait();
#pragma unroll
        for (int e = 0; e < 16; ++e) {
          const float v = __uint_as_float(r[e]);
          acc[k * 16 + e] = fmaf(fmaf(fabsf(v), c, v), m, acc[k * 16 + e]);
        }
      }
    } else if (MODE == 8) {
      uint32_t r[48];
#pragma unroll
      for (int k = 0; k < 3; ++k) ld16(tl + q * 48 + k * 16 + (it & 1) * 256, r + 16 * k);
      ldwait();
#pragma unroll
      for (int e = 0; e < 48; ++e) {
        const float v = __uint_as_float(r[e]);
        acc[e] = fmaf(fmaf(fabsf(v), c, v), m, acc[e]);
      }
    } else {  // E1 variants: 40 columns -> bf16 swizzled tile
#pragma unroll
      for (int c0 = 0; c0 < 40; c0 += 16) {
        const int n = (40 - c0) >= 16 ? 16 : 8;
        uint32_t r[16];
        if (n == 16) ld16(tl + q * 40 + c0, r);
        else {
          asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                       : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                       : "r"(tl + q * 40 + c0));
        }
        ldwait();
        uint32_t p[8];
#pragma unroll
        for (int e = 0; e < 16; e += 2) {
          if (e < n) {
            float a = __uint_as_float(r[e]), b = __uint_as_float(r[e + 1]);
            if (MODE == 5) {
              a = fmaxf(a, alpha * a); b = fmaxf(b, alpha * b);
              p[e / 2] = pack_bf16(a, b);
            } else if (MODE == 6) {
              a = fmaf(fabsf(a), c, a); b = fmaf(fabsf(b), c, b);
              p[e / 2] = pack_bf16(a, b);
            } else {
              __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
              const __nv_bfloat162 al = __float2bfloat162_rn(alpha);
              h = __hmax2(h, __hmul2(h, al));
              p[e / 2] = *reinterpret_cast<uint32_t*>(&h);
            }
          }
        }
        const uint32_t k = q * 40 + c0;
        uint32_t addr = sbase + (k >> 6) * 16384 + row * 128 + ((((k & 63) >> 3) ^ (row & 7)) << 4);
        asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(p[0]), "r"(p[1]), "r"(p[2]), "r"(p[3]));
        if (n == 16) {
          const uint32_t k2 = k + 8;
          addr = sbase + (k2 >> 6) * 16384 + row * 128 + ((((k2 & 63) >> 3) ^ (row & 7)) << 4);
          asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(p[4]), "r"(p[5]), "r"(p[6]), "r"(p[7]));
        }
      }
    }
  }
  const long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 48; ++i) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}

template <int MODE>
void run(const char* name, int cols_per_thread_iter) {
  float* out;
  long long* clk;
  cudaMalloc(&out, 148 * 512 * 4);
  cudaMalloc(&clk, 148 * 8);
  cudaFuncSetAttribute(bench<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  const int iters = 2000;
  for (int threads : {128, 256, 512}) {
    bench<MODE><<<1, threads, 64 * 1024>>>(iters, 0.2f, 0.7f, out, clk);
    cudaDeviceSynchronize();
    bench<MODE><<<1, threads, 64 * 1024>>>(iters, 0.2f, 0.7f, out, clk);
    cudaError_t e = cudaDeviceSynchronize();
    long long c = 0;
    cudaMemcpy(&c, clk, 8, cudaMemcpyDeviceToHost);
    const double per_it = (double)c / iters;
    const double bytes = (double)threads * cols_per_thread_iter * 4;
    printf("%-44s warps=%2d  %8.1f clk/iter  %6.1f B/clk/SM (TMEM read)  %s\n", name, threads / 32, per_it,
           bytes / per_it, e == cudaSuccess ? "" : cudaGetErrorString(e));
  }
  cudaFree(out);
  cudaFree(clk);
}

int main() {
  run<0>("ld x16 + wait, x3 (48 cols)", 48);
  run<1>("3 x ld x16, one wait (48 cols)", 48);
  run<2>("ld x32 + wait (32 cols)", 32);
  run<3>("E2 v1: fmul+fmax+ffma (48 cols)", 48);
  run<4>("E2 v2: ffma|v| + ffma (48 cols)", 48);
  run<8>("E2 v2, 3 lds one wait (48 cols)", 48);
  run<5>("E1 v1: fmul+fmax+cvt+sts (40 cols)", 40);
  run<6>("E1 v2: ffma|v|+cvt+sts (40 cols)", 40);
  run<7>("E1 v3: cvt+hmul2+hmax2+sts (40 cols)", 40);
  return 0;
}
