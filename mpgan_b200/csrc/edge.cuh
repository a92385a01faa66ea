// Argument block shared by the edge-network kernels (generic SIMT and tcgen05).
#pragma once
#include <type_traits>

#include "common.cuh"

namespace mpg {

struct EdgeArgs {
  int B, N, F, H0, H1, H2;
  // factorised first layer (node level): P = x*Wa^T + b0, Q = x*Wb^T   [B*N, H0]
  const float* P;
  const float* Q;
  // p_tiled: P and dP are stored per 128-row tile as [column group k/4][row in tile][4] (the order in which the
  // tcgen05 kernels' threads -- one per row -- read them: a warp's 32 rows x 16 bytes are contiguous); only set when
  // both ends (pq kernels and tcgen05 edge kernels) run, with both buffers rounded up to whole tiles.  Q/dQ stay
  // row-major (they are copied a row at a time).
  int p_tiled;
  // set together with p_tiled: P / Q have NOT been computed yet -- the tcgen05 launcher's set-up kernel does it (next to
  // the weight images and the work list, one launch) from x, W0 (= Wef - 2F, row stride ldwef) and b0
  int pq_deferred;
  const float* b0;
  // Receiver compaction (tcgen05 path, opt-in: mpg_edge_set_compaction): tiles are built from the rows whose mask is
  // non-zero only, so padded particles cost nothing as RECEIVERS either (they already cost nothing as senders).  Every
  // global tensor keeps its padded [B*N, .] layout; `cmap` (device, written by mpg_compact_map) says which padded row
  // each of a tile's 128 lanes stands for.  Rows that are in no tile keep agg = 0 / dx = 0: exact for a network whose
  // padded particles are dropped downstream (the discriminator: masked as senders, multiplied by the mask at the
  // pooling, mpgan/model.py:810-822,881-884) -- not for the generator, whose padded rows are part of its output.
  //   cmap[0] = number of tiles, cmap[2 + t] = first jet of tile t, cmap[2 + T + t] = jets it spans (<= 10),
  //   cmap[2 + 2T + 128 t + lane] = padded row (b*N + i) or -1, T = ctiles_max
  const int* cmap;
  int ctiles_max;
  // optional pair features (pos_diffs): ef_mode bit0 = distance column, bit1 = difference columns
  const float* x;        // [B*N, F] node features, row stride ldx
  int ldx;
  const float* Wef;      // &W0[0][2F], row stride ldwef
  int ldwef, n_ef, nd, ef_mode;
  const float* W1;       // [H1, H0]  (reference layout)
  const float* W2;       // [H2, H1]
  const float* W1t;      // [H0, H1]  (transposed copy, workspace)
  const float* W2t;      // [H1, H2]
  const float* b1;
  const float* b2;
  const float* mask;     // [B*N] multiplier on the sender axis, or null
  float* agg;            // [B*N, H2]
  float alpha, out_scale;
  DropCfg drop;
  // backward
  const float* dagg;     // [B*N, H2]
  float *dW1, *db1, *dW2, *db2;   // accumulated (atomics)
  float *dP, *dQ;        // [B*N, H0]; dQ must be zeroed by the caller
  float* dWef;           // &dW0[0][2F], row stride ldwef (accumulated)
  float* dx_ef;          // [B*N, F] zeroed by the caller
  // ---- generic (fp32 SIMT) kernel only --------------------------------------------------------------------------
  // kNN message passing (mpgan/model.py:319-381): receiver i sends over its K listed neighbours instead of all N
  // particles; the distance feature is taken to the sender scaled by (1 - 1e4) * mask + 1e4 (:336-340)
  const int* nbr;        // [B*N, K] sender indices inside the jet, or null (fully connected)
  int K;                 // senders per receiver (N when nbr is null)
  int knn_scale;         // scale masked senders in the distance feature (kNN with a mask)
  float* dmask;          // optional [B*N]: d/d(mask), accumulated with atomics (zeroed by the caller)
  // conditioning columns of the edge network (clabels / mask_fne_np, mpgan/model.py:247-253): pair row r of the
  // [B*N*K, .] edge input gets cond[r % B] (the reference's .repeat quirk), i.e. the first layer gains
  // Lc[r % B] = W0c cond[r % B].  Lc [B, H0]; dLc [B, H0] (zeroed by the caller) receives its gradient.
  const float* Lc;
  float* dLc;
  // second-order ("tangent") mode of the backward kernel, see mpg_edge_bwd2: Pt/Qt = first layer applied to the
  // direction u (no bias); tagg [B*N, H2] and gmask [B*N] (zeroed by the caller) are outputs; dW1/dW2 receive the
  // second-order weight gradients, dP/dQ the ordinary first-order ones
  const float* Pt;
  const float* Qt;
  float* tagg;
  float* gmask;
};

size_t edge_generic_smem(const EdgeArgs& a, bool bwd, bool tangent = false);
int launch_edge_generic(const EdgeArgs& a, bool bwd, cudaStream_t stream);
int launch_edge_generic_tangent(const EdgeArgs& a, cudaStream_t stream);
int launch_transpose(const float* in, int rows, int cols, float* out, cudaStream_t stream);

// node-level ends of the factorised first layer (edge_node.cu)
bool pq_supported(int F, int H0);
int launch_pq_fwd(const float* x, int ldx, const float* W0, int ldw, const float* b0, float* P, float* Q, int BN,
                  int F, int H0, cudaStream_t stream, bool p_tiled = false);
int launch_pq_bwd(const float* dP, const float* dQ, const float* x, int ldx, const float* W0, int ldw, float* dx,
                  int lddx, float* dW0, float* db0, int BN, int F, int H0, cudaStream_t stream, bool p_tiled = false,
                  bool tf32 = false,    // tf32: TF32 tensor-core products (precision 1)
                  const int* cmap = nullptr, int ctiles_max = 0);   // receiver compaction (EdgeArgs::cmap)
// compaction helpers (host + device)
__host__ __device__ inline int compact_tiles_max(long long B, long long N) {   // every jet is max(n, 15) <= max(N, 15) wide
  return (int)((B * (N > 15 ? N : 15) + 127) / 128);
}
__host__ __device__ inline size_t compact_map_ints(long long B, long long N) {
  return 2 + (size_t)compact_tiles_max(B, N) * (2 + 128);
}
// element (row r, column k) of a tiled P / dP buffer (EdgeArgs::p_tiled), H0 columns, 128-row tiles
__host__ __device__ inline size_t p_tiled_index(size_t r, int k, int H0) {
  return (r >> 7) * 128 * (size_t)H0 + ((size_t)(k >> 2) * 128 + (r & 127)) * 4 + (k & 3);
}

// tcgen05 path (edge_tc.cu): default architecture only
int edge_tc_features();
void edge_tc_arm_probe(int id, cudaEvent_t e0, cudaEvent_t e1);
bool edge_tc_supported(const EdgeArgs& a);
// workspace in two parts: `persist` (weight images + work list: written by the forward, reusable by the backward
// of the same call -- reuse = true skips the kernels that build them) and `scratch` (backward only: sign bits,
// weight-gradient slabs)
size_t edge_tc_persist_bytes(int B, int N, int H0, int H1, int H2);
size_t edge_tc_scratch_bytes(int B, int N, int H0, int H1, int H2);
int launch_edge_tc_fwd(const EdgeArgs& a, void* persist, cudaStream_t stream);
int launch_edge_tc_bwd(const EdgeArgs& a, void* persist, void* scratch, bool reuse, cudaStream_t stream);

}  // namespace mpg
