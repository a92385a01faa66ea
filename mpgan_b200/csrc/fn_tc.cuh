// Fused tcgen05 node network (fn_tc.cu): three chained TF32 GEMMs per 128-row tile, forward and
// input-gradient backward.  See fn_tc.cu for the data flow.
#pragma once
#include "common.cuh"

namespace mpg {

struct FnImageJob {
  const float* W;     // source matrix [R, C], row stride ldw
  int ldw, R, C;
  int transposed;     // image element (n, k) = W[k][n] instead of W[n][k]
  int rows, kblocks;  // image: `rows` (multiple of 16) x 32*kblocks, zero padded
  uint8_t* dst;
};
struct FnImageJobs { FnImageJob job[3]; };

struct FnTcArgs {
  int M;                        // rows (particles)
  int n[3], kmma[3];            // per GEMM: MMA N (padded) and number of K = 8 steps   (filled by launch_fn_tc)
  const uint8_t* img[3];        // weight images                                        (filled by launch_fn_tc)
  // first A operand [a | b]: forward a = agg, b = x; backward a = dout, Kb = 0
  const float* a; int lda, Ka;
  const float* b; int ldb, Kb;
  const float* bias[3];         // forward
  const float* ysave[2];        // backward: {y1, y0} (outputs of the layer whose derivative epilogue l applies)
  float* out01[2];              // forward {y0, y1}; backward {dz1, dz0}; contiguous [M, n[l]]
  float* dz2;                   // backward with dropout: dout * drop' [M, NO]
  float* dbias[3];              // backward: bias gradients of layer 0, 1, 2 (accumulated; NULL = skip)
  // last GEMM's output columns [0, Na) -> outa, [Na, Na+Nb) -> outb   (forward: out, Nb = 0; backward: da, db)
  float* outa; int ldoa, Na;
  float* outb; int ldob, Nb;
  float alpha;
  DropCfg drop;                 // p in {0, 0.5}
  uint32_t stream[3];           // RNG stream of layer 0, 1, 2 (LinearNet: 16 + i)
  int vec_a, vec_b, vec_oa, vec_ob;   // 16-byte access allowed                        (filled by launch_fn_tc)
};

bool fn_tc_supported(int Ka, int Kb, int H1, int H2, int NO, float p);
size_t fn_tc_workspace_bytes(int Ka, int Kb, int H1, int H2, int NO);
// w0 [H1, Ka+Kb], w1 [H2, H1], w2 [NO, H2] (reference layout, contiguous); ws >= fn_tc_workspace_bytes
int launch_fn_tc(FnTcArgs t, bool bwd, const float* w0, const float* w1, const float* w2, int H1, int H2, int NO,
                 void* ws, cudaStream_t stream);

// weight gradients dW_l += dz_l^T [ina_l | inb_l] for up to three products over the same M rows in one launch
// (fn_dw_kernel): the node network's three layers, or the two halves of the factorised first edge layer
struct FnDwArgs {
  int M, T;                        // rows; 128-row tiles (filled by launch_fn_dw)
  int nslots, items_per_cta;       // products in this launch (1..3); 128-row items a CTA should own (>= 1)
  const float* dz[3]; int na[3];   // dz_l [M, na] contiguous
  const float* ina[3]; int lda[3], ka[3];
  const float* inb[3]; int ldb[3], kb[3];
  float* dw[3]; int lddw[3];       // dW_l [na, ka + kb], row stride lddw (accumulated)
  int vec_dz[3], vec_a[3], vec_b[3], vec_dw[3];   // filled by launch_fn_dw
};
int launch_fn_dw(FnDwArgs t, cudaStream_t stream);

}  // namespace mpg
