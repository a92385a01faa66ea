// extern "C" boundary of libmpgan_b200.so (see include/mpgan_b200.h).
#include <stdarg.h>
#include <string.h>

#include "../../include/mpgan_b200.h"
#include "edge.cuh"
#include "extra.cuh"
#include "fn_tc.cuh"
#include "gapt.cuh"
#include "gemm.cuh"
#include "mab.cuh"
#include "misc.cuh"
#include "peer.cuh"

namespace mpg {
static thread_local char g_err[512] = "";
unsigned long long g_launch_count = 0;
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

namespace {

__global__ void add_strided_kernel(float* __restrict__ dst, int ldd, const float* __restrict__ src, int lds,
                                   int rows, int cols) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx < rows * cols) {
    const int r = idx / cols, c = idx % cols;
    dst[(size_t)r * ldd + c] += src[(size_t)r * lds + c];
  }
}

// one helper stream + fork/join events per device, created on first use (never destroyed: process lifetime)
struct SideStream {
  cudaStream_t stream = nullptr;
  cudaEvent_t fork = nullptr, join = nullptr;
};
SideStream* side_stream() {
  static SideStream per_dev[64];
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
  SideStream& ss = per_dev[dev];
  if (ss.stream == nullptr) {
    if (cudaStreamCreateWithFlags(&ss.stream, cudaStreamNonBlocking) != cudaSuccess) return nullptr;
    if (cudaEventCreateWithFlags(&ss.fork, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&ss.join, cudaEventDisableTiming) != cudaSuccess) {
      ss.stream = nullptr;
      return nullptr;
    }
  }
  return &ss;
}

// Workspace of one fused edge call.  The first `persist` bytes (P, Q, the transposed / swizzled weight copies,
// the work list) are produced by the forward and depend only on (x, weights, mask): the backward of the same
// call can read them from the forward's buffer instead of recomputing them (mpg_edge_bwd_saved).
struct EdgeWs {
  float *P, *Q, *W1t, *W2t, *dP, *dQ, *dxef, *wg_scratch;
  void *tc, *tc_scratch;
  size_t persist, total;
};

size_t align_up(size_t v) { return (v + 255) & ~(size_t)255; }

EdgeWs carve_edge_ws(void* base, int B, int N, int F, int H0, int H1, int H2) {
  EdgeWs w;
  size_t off = 0;
  char* p = reinterpret_cast<char*>(base);
  auto take = [&](size_t floats) {
    float* r = reinterpret_cast<float*>(p + off);
    off += align_up(floats * sizeof(float));
    return r;
  };
  const size_t BN = (size_t)B * N;
  // P / dP may be stored per 128-row tile (EdgeArgs::p_tiled), of the compacted tile space when receivers are compacted
  const size_t BNT = (size_t)compact_tiles_max(B, N) * 128;
  w.P = take(BNT * H0);
  w.Q = take(BN * H0);
  w.W1t = take((size_t)H0 * H1);
  w.W2t = take((size_t)H1 * H2);
  w.tc = p + off;
  off += align_up(edge_tc_persist_bytes(B, N, H0, H1, H2));
  w.persist = off;
  w.dP = take(BNT * H0);    // dP and dQ are adjacent: one memset clears both
  w.dQ = take(BN * H0);
  w.dxef = take(BN * F);
  // sink for the weight gradients of a dx-only backward on the generic path (never read)
  // (first-layer rows are 2F + n_ef wide, n_ef <= F + 1)
  w.wg_scratch = take((size_t)H0 * (3 * F + 1) + H0 + (size_t)H1 * H0 + H1 + (size_t)H2 * H1 + H2);
  w.tc_scratch = p + off;
  off += align_up(edge_tc_scratch_bytes(B, N, H0, H1, H2));
  w.total = off;
  return w;
}

// optional operands of the fp32 kernels: neighbour list (kNN) and the mask gradient
struct EdgeExtra {
  const int* nbr = nullptr;
  int K = 0;
  int knn_scale = 0;
  float* dmask = nullptr;
  const float* lc = nullptr;
  float* dlc = nullptr;
  bool fp32_only = false;   // entry points that always run the fp32 kernels
  bool any() const { return fp32_only || nbr != nullptr || dmask != nullptr || lc != nullptr; }
};

// receiver compaction map of this thread's next edge calls (mpg_edge_set_compaction); null: off
static thread_local const int* g_cmap = nullptr;

int edge_common(EdgeArgs& a, EdgeWs& w, const float* x, int ldx, const float* mask, const float* w0,
                const float* b0, const float* w1, const float* b1, const float* w2, const float* b2, int B, int N,
                int F, int H0, int H1, int H2, int ef_mode, int nd, int mean, float alpha, float p_drop,
                uint64_t seed, const uint64_t* seed_dev, int precision, void* workspace, size_t workspace_bytes, bool* use_tc,
                cudaStream_t s, bool fwd_only = false, const void* saved = nullptr, size_t saved_bytes = 0) {
  MPG_CHECK(B > 0 && N > 0 && F > 0, "bad edge problem size B=%d N=%d F=%d", B, N, F);
  MPG_CHECK(ef_mode >= 0 && ef_mode <= 3 && (ef_mode == 0 || (nd > 0 && nd <= F)), "bad ef_mode/nd");
  MPG_CHECK(p_drop >= 0.f && p_drop < 1.f, "dropout p must be in [0,1)");
  w = carve_edge_ws(workspace, B, N, F, H0, H1, H2);
  const size_t need = fwd_only ? w.persist : w.total;
  MPG_CHECK(workspace != nullptr && workspace_bytes >= need, "edge workspace too small: need %zu bytes, got %zu",
            need, workspace_bytes);
  if (saved != nullptr) {   // P, Q, weight copies and work list come from the forward's workspace
    MPG_CHECK(saved_bytes >= w.persist, "saved forward workspace too small: need %zu bytes, got %zu", w.persist,
              saved_bytes);
    const EdgeWs f = carve_edge_ws(const_cast<void*>(saved), B, N, F, H0, H1, H2);
    w.P = f.P; w.Q = f.Q; w.W1t = f.W1t; w.W2t = f.W2t; w.tc = f.tc;
  }
  memset(&a, 0, sizeof(a));
  a.B = B; a.N = N; a.F = F; a.H0 = H0; a.H1 = H1; a.H2 = H2;
  a.n_ef = ((ef_mode & 2) ? nd : 0) + (ef_mode & 1);
  a.nd = ef_mode ? nd : 0;
  a.ef_mode = ef_mode;
  a.ldwef = 2 * F + a.n_ef;
  a.Wef = w0 + 2 * F;
  a.x = x; a.ldx = ldx;
  a.P = w.P; a.Q = w.Q;
  a.W1 = w1; a.W2 = w2; a.W1t = w.W1t; a.W2t = w.W2t; a.b1 = b1; a.b2 = b2;
  a.mask = mask;
  a.alpha = alpha;
  a.out_scale = mean ? 1.f / (float)N : 1.f;
  a.drop = make_drop(p_drop, seed, seed_dev);
  const bool precise = precision == 0;
  *use_tc = !precise && edge_tc_supported(a);
  // both ends of P / dP are the pq kernels and the tcgen05 kernels: tile-major storage (the backward takes the same
  // decision from the same arguments, so a saved forward workspace is read the way it was written)
  a.p_tiled = (*use_tc && (edge_tc_features() & 2) && pq_supported(F, H0)) ? 1 : 0;
  if (saved != nullptr) {
    if (a.p_tiled && g_cmap != nullptr) {   // the forward that filled `saved` ran compacted: same map
      a.cmap = g_cmap;
      a.ctiles_max = compact_tiles_max(B, N);
    }
    return 0;
  }
  // factorised first layer: W0 [x_i ; x_j ; ef] = Wa x_i + Wb x_j + Wef ef   (node-level GEMMs)
  a.pq_deferred = a.p_tiled;
  a.b0 = b0;
  if (a.p_tiled && g_cmap != nullptr) {   // receiver compaction requested for this thread's edge calls (tcgen05 path only)
    a.cmap = g_cmap;
    a.ctiles_max = compact_tiles_max(B, N);
  }
  if (a.pq_deferred) {
    // computed by the tcgen05 launcher's set-up kernel
  } else if (pq_supported(F, H0)) {
    if (launch_pq_fwd(x, ldx, w0, a.ldwef, b0, w.P, w.Q, B * N, F, H0, s, false)) return 1;
  } else {
    GemmEpi e;
    e.bias = b0;
    if (launch_gemm(true, true, true, x, ldx, w0, a.ldwef, w.P, H0, B * N, H0, F, e, 1, s)) return 1;
    GemmEpi e2;
    if (launch_gemm(true, true, true, x, ldx, w0 + F, a.ldwef, w.Q, H0, B * N, H0, F, e2, 1, s)) return 1;
  }
  if (!*use_tc) {
    if (launch_transpose(w1, H1, H0, w.W1t, s)) return 1;
    if (launch_transpose(w2, H2, H1, w.W2t, s)) return 1;
  }
  return 0;
}

}  // namespace
}  // namespace mpg

using namespace mpg;

extern "C" {

int mpg_version(void) { return 100; }
const char* mpg_last_error(void) { return g_err; }
int mpg_features(void) { return edge_tc_features() | 4; }
unsigned long long mpg_launch_count(void) { return g_launch_count; }
void mpg_probe(int kernel_id, void* ev_start, void* ev_stop) {
  edge_tc_arm_probe(kernel_id, (cudaEvent_t)ev_start, (cudaEvent_t)ev_stop);
}

int mpg_linear_fwd(const float* x, int ldx, const float* w, const float* b, float* y, int M, int K, int N, int act,
                   float alpha, float p_drop, uint64_t seed, const uint64_t* seed_dev, uint32_t rng_stream, int precision,
                   void* stream) {
  MPG_CHECK(M >= 0 && K > 0 && N > 0, "bad linear size M=%d K=%d N=%d", M, K, N);
  MPG_CHECK(p_drop >= 0.f && p_drop < 1.f, "dropout p must be in [0,1)");
  GemmEpi e;
  e.bias = b;
  e.act = act;
  e.alpha = alpha;
  e.drop = p_drop > 0.f;
  e.dc = make_drop(p_drop, seed, seed_dev);
  e.stream = rng_stream;
  return launch_gemm(true, true, precision == 0, x, ldx, w, K, y, N, M, N, K, e, 1, (cudaStream_t)stream);
}

int mpg_linear_bwd(const float* dy, const float* y, const float* x, int ldx, const float* w, float* dz, float* dx,
                   int lddx, int dx_accumulate, float* dw, float* db, int M, int K, int N, int act, float alpha,
                   float p_drop, uint64_t seed, const uint64_t* seed_dev, uint32_t rng_stream, int precision,
                   void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  MPG_CHECK(M >= 0 && K > 0 && N > 0, "bad linear size M=%d K=%d N=%d", M, K, N);
  const bool precise = precision == 0;
  const float* g = dy;
  if (act || p_drop > 0.f) {
    MPG_CHECK(dz != nullptr && y != nullptr, "linear_bwd needs y and a dz scratch buffer");
    if (launch_act_bwd(dy, y, dz, M, N, act, alpha, make_drop(p_drop, seed, seed_dev), rng_stream, s)) return 1;
    g = dz;
  }
  // The input-gradient GEMM and the weight/bias-gradient kernels are independent and each too small to fill
  // the GPU (2-4 CTAs per SM, latency-bound): fork the weight side onto a second stream and join before
  // returning.  Stream-ordered events only, so the fork/join is captured into CUDA graphs as parallel branches.
  SideStream* side = (dx != nullptr && (dw != nullptr || db != nullptr)) ? side_stream() : nullptr;
  cudaStream_t sw = s;
  if (side != nullptr) {
    MPG_CUDA(cudaEventRecord(side->fork, s));
    MPG_CUDA(cudaStreamWaitEvent(side->stream, side->fork, 0));
    sw = side->stream;
  }
  if (dw != nullptr) {
    GemmEpi e;
    e.accumulate = 1;
    const int tiles = cdiv(N, 64) * cdiv(K, 64);
    int split = tiles >= 148 ? 1 : (2 * 148) / tiles;
    if (split > cdiv(M, 128)) split = cdiv(M, 128);
    if (launch_gemm(false, false, precise, g, N, x, ldx, dw, K, N, K, M, e, split < 1 ? 1 : split, sw)) return 1;
  }
  if (db != nullptr)
    if (launch_colsum(g, N, M, N, db, sw)) return 1;
  if (dx != nullptr) {
    GemmEpi e;
    e.accumulate = dx_accumulate;
    if (launch_gemm(true, false, precise, g, N, w, K, dx, lddx, M, K, N, e, 1, s)) return 1;
  }
  if (side != nullptr) {
    MPG_CUDA(cudaEventRecord(side->join, side->stream));
    MPG_CUDA(cudaStreamWaitEvent(s, side->join, 0));
  }
  return 0;
}

int mpg_fn_supported(int Ka, int Kb, int H1, int H2, int NO, float p_drop) {
  return fn_tc_supported(Ka, Kb, H1, H2, NO, p_drop) ? 1 : 0;
}
size_t mpg_fn_workspace_bytes(int Ka, int Kb, int H1, int H2, int NO) {
  return fn_tc_workspace_bytes(Ka, Kb, H1, H2, NO);
}

int mpg_fn_fwd(const float* a, int lda, int Ka, const float* b, int ldb, int Kb, int M, const float* w0,
               const float* b0, const float* w1, const float* b1, const float* w2, const float* b2, int H1, int H2,
               int NO, float alpha, float p_drop, uint64_t seed, const uint64_t* seed_dev, void* workspace,
               size_t workspace_bytes, float* y0, float* y1, float* out, void* stream) {
  MPG_CHECK(fn_tc_supported(Ka, Kb, H1, H2, NO, p_drop), "fn: unsupported shape Ka=%d Kb=%d H1=%d H2=%d NO=%d p=%g",
            Ka, Kb, H1, H2, NO, (double)p_drop);
  MPG_CHECK(workspace != nullptr && workspace_bytes >= fn_tc_workspace_bytes(Ka, Kb, H1, H2, NO),
            "fn workspace too small");
  if (M <= 0) return 0;
  FnTcArgs t;
  memset(&t, 0, sizeof(t));
  t.M = M;
  t.a = a; t.lda = lda; t.Ka = Ka; t.b = b; t.ldb = ldb; t.Kb = Kb;
  t.bias[0] = b0; t.bias[1] = b1; t.bias[2] = b2;
  t.out01[0] = y0; t.out01[1] = y1;
  t.outa = out; t.ldoa = NO; t.Na = NO; t.outb = nullptr; t.ldob = 0; t.Nb = 0;
  t.alpha = alpha;
  t.drop = make_drop(p_drop, seed, seed_dev);
  t.stream[0] = 16; t.stream[1] = 17; t.stream[2] = 18;
  return launch_fn_tc(t, false, w0, w1, w2, H1, H2, NO, workspace, (cudaStream_t)stream);
}

int mpg_fn_bwd(const float* dout, const float* y0, const float* y1, const float* a, int lda, int Ka, const float* b,
               int ldb, int Kb, int M, const float* w0, const float* w1, const float* w2, int H1, int H2, int NO,
               float alpha, float p_drop, uint64_t seed, const uint64_t* seed_dev, void* workspace,
               size_t workspace_bytes, float* dz0, float* dz1, float* dz2, float* da, float* db, float* dw0,
               float* db0, float* dw1, float* db1, float* dw2, float* db2, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  MPG_CHECK(fn_tc_supported(Ka, Kb, H1, H2, NO, p_drop), "fn: unsupported shape Ka=%d Kb=%d H1=%d H2=%d NO=%d p=%g",
            Ka, Kb, H1, H2, NO, (double)p_drop);
  MPG_CHECK(workspace != nullptr && workspace_bytes >= fn_tc_workspace_bytes(Ka, Kb, H1, H2, NO),
            "fn workspace too small");
  MPG_CHECK(dz0 != nullptr && dz1 != nullptr && (p_drop == 0.f || dz2 != nullptr), "fn_bwd needs its dz scratch buffers");
  if (M <= 0) return 0;
  FnTcArgs t;
  memset(&t, 0, sizeof(t));
  t.M = M;
  t.a = dout; t.lda = NO; t.Ka = NO; t.b = nullptr; t.ldb = 0; t.Kb = 0;
  t.ysave[0] = y1; t.ysave[1] = y0;
  t.out01[0] = dz1; t.out01[1] = dz0;
  t.dz2 = dz2;
  t.outa = da; t.ldoa = Ka; t.Na = Ka; t.outb = db; t.ldob = Kb; t.Nb = Kb;
  t.alpha = alpha;
  t.drop = make_drop(p_drop, seed, seed_dev);
  t.stream[0] = 16; t.stream[1] = 17; t.stream[2] = 18;
  if (dw0 != nullptr) { t.dbias[0] = db0; t.dbias[1] = db1; t.dbias[2] = db2; }   // column sums in the epilogues
  if (launch_fn_tc(t, true, w0, w1, w2, H1, H2, NO, workspace, s)) return 1;
  if (dw0 == nullptr) return 0;
  // weight / bias gradients: dW_l += dz_l^T (input of layer l), db_l += colsum(dz_l) on the side stream (the
  // caller's next kernels need only da / db)
  SideStream* side = side_stream();
  cudaStream_t sw = s;
  if (side != nullptr) {
    MPG_CUDA(cudaEventRecord(side->fork, s));
    MPG_CUDA(cudaStreamWaitEvent(side->stream, side->fork, 0));
    sw = side->stream;
  }
  const float* g2 = p_drop > 0.f ? dz2 : dout;
  {
    FnDwArgs d;
    memset(&d, 0, sizeof(d));
    d.M = M;
    d.nslots = 3;
    d.items_per_cta = 1;
    // slot order = processing order: the two big layers first, the narrow last layer at the end
    d.dz[0] = dz1; d.na[0] = H2; d.ina[0] = y0; d.lda[0] = H1; d.ka[0] = H1; d.dw[0] = dw1; d.lddw[0] = H1;
    d.dz[1] = dz0; d.na[1] = H1; d.ina[1] = a; d.lda[1] = lda; d.ka[1] = Ka; d.inb[1] = b; d.ldb[1] = ldb; d.kb[1] = Kb;
    d.dw[1] = dw0; d.lddw[1] = Ka + Kb;
    d.dz[2] = g2; d.na[2] = NO; d.ina[2] = y1; d.lda[2] = H2; d.ka[2] = H2; d.dw[2] = dw2; d.lddw[2] = H2;
    if (launch_fn_dw(d, sw)) return 1;
  }
  if (side != nullptr) {
    MPG_CUDA(cudaEventRecord(side->join, side->stream));
    MPG_CUDA(cudaStreamWaitEvent(s, side->join, 0));
  }
  return 0;
}

size_t mpg_edge_workspace_bytes(int B, int N, int F, int H0, int H1, int H2) {
  return carve_edge_ws(nullptr, B, N, F, H0, H1, H2).total;
}
size_t mpg_edge_fwd_workspace_bytes(int B, int N, int F, int H0, int H1, int H2) {
  return carve_edge_ws(nullptr, B, N, F, H0, H1, H2).persist;
}

static int edge_fwd_impl(const EdgeExtra& ex, const float* x, int ldx, const float* mask, const float* w0,
                         const float* b0, const float* w1, const float* b1, const float* w2, const float* b2, int B, int N,
                         int F, int H0, int H1, int H2, int ef_mode, int nd, int mean, float alpha, float p_drop,
                         uint64_t seed, const uint64_t* seed_dev, int precision, void* workspace, size_t workspace_bytes,
                         float* agg, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  EdgeArgs a;
  EdgeWs w;
  bool use_tc = false;
  if (edge_common(a, w, x, ldx, mask, w0, b0, w1, b1, w2, b2, B, N, F, H0, H1, H2, ef_mode, nd, mean, alpha,
                  p_drop, seed, seed_dev, ex.any() ? 0 : precision, workspace, workspace_bytes, &use_tc, s, true))
    return 1;
  a.agg = agg;
  a.Lc = ex.lc;
  if (ex.nbr != nullptr) {
    MPG_CHECK(ex.K > 0 && ex.K <= N, "edge_nbr: need 0 < K <= N (K = %d, N = %d)", ex.K, N);
    a.nbr = ex.nbr; a.K = ex.K; a.knn_scale = ex.knn_scale;
    if (mean) a.out_scale = 1.f / (float)ex.K;   // torch.mean over the num_knn axis (model.py:267)
  }
  if (use_tc) return launch_edge_tc_fwd(a, w.tc, s);
  return launch_edge_generic(a, false, s);
}

int mpg_edge_fwd(const float* x, int ldx, const float* mask, const float* w0, const float* b0, const float* w1,
                 const float* b1, const float* w2, const float* b2, int B, int N, int F, int H0, int H1, int H2,
                 int ef_mode, int nd, int mean, float alpha, float p_drop, uint64_t seed, const uint64_t* seed_dev,
                 int precision, void* workspace, size_t workspace_bytes, float* agg, void* stream) {
  return edge_fwd_impl(EdgeExtra(), x, ldx, mask, w0, b0, w1, b1, w2, b2, B, N, F, H0, H1, H2, ef_mode, nd, mean, alpha,
                       p_drop, seed, seed_dev, precision, workspace, workspace_bytes, agg, stream);
}

int mpg_edge_nbr_fwd(const int* nbr, int K, int knn_scale, const float* lc, const float* x, int ldx, const float* mask,
                     const float* w0, const float* b0, const float* w1, const float* b1, const float* w2, const float* b2,
                     int B, int N, int F, int H0, int H1, int H2, int ef_mode, int nd, int mean, float alpha, float p_drop,
                     uint64_t seed, const uint64_t* seed_dev, void* workspace, size_t workspace_bytes, float* agg,
                     void* stream) {
  EdgeExtra ex;
  ex.fp32_only = true;
  ex.nbr = nbr; ex.K = K; ex.knn_scale = knn_scale; ex.lc = lc;
  return edge_fwd_impl(ex, x, ldx, mask, w0, b0, w1, b1, w2, b2, B, N, F, H0, H1, H2, ef_mode, nd, mean, alpha, p_drop,
                       seed, seed_dev, 0, workspace, workspace_bytes, agg, stream);
}

static int edge_bwd_impl(const EdgeExtra& ex, const void* saved, size_t saved_bytes, const float* x, int ldx, const float* mask,
                         const float* w0, const float* b0, const float* w1, const float* b1, const float* w2,
                         const float* b2, int B, int N, int F, int H0, int H1, int H2, int ef_mode, int nd, int mean,
                         float alpha, float p_drop, uint64_t seed, const uint64_t* seed_dev, int precision,
                         void* workspace, size_t workspace_bytes, const float* dagg, float* dx, int lddx, float* dw0,
                         float* db0, float* dw1, float* db1, float* dw2, float* db2, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  EdgeArgs a;
  EdgeWs w;
  bool use_tc = false;
  if (edge_common(a, w, x, ldx, mask, w0, b0, w1, b1, w2, b2, B, N, F, H0, H1, H2, ef_mode, nd, mean, alpha,
                  p_drop, seed, seed_dev, ex.any() ? 0 : precision, workspace, workspace_bytes, &use_tc, s, false, saved,
                  saved_bytes))
    return 1;
  MPG_CHECK(dagg && dx, "edge_bwd: null gradient pointer");
  if (ex.nbr != nullptr) {
    MPG_CHECK(ex.K > 0 && ex.K <= N, "edge_nbr: need 0 < K <= N (K = %d, N = %d)", ex.K, N);
    a.nbr = ex.nbr; a.K = ex.K; a.knn_scale = ex.knn_scale;
    if (mean) a.out_scale = 1.f / (float)ex.K;
  }
  a.Lc = ex.lc;
  if (ex.lc != nullptr && ex.dlc != nullptr) {
    a.dLc = ex.dlc;
    MPG_CUDA(cudaMemsetAsync(ex.dlc, 0, (size_t)B * H0 * sizeof(float), s));
  }
  if (ex.dmask != nullptr) {
    MPG_CHECK(mask != nullptr, "edge_bwd: a mask gradient needs a mask");
    a.dmask = ex.dmask;
    MPG_CUDA(cudaMemsetAsync(ex.dmask, 0, (size_t)B * N * sizeof(float), s));
  }
  // all six weight-gradient pointers null = input gradient only (train_G back-propagates through a frozen D)
  const bool dx_only = !dw0 && !db0 && !dw1 && !db1 && !dw2 && !db2;
  MPG_CHECK(dx_only || (dw0 && db0 && dw1 && db1 && dw2 && db2), "edge_bwd: pass all six weight gradients or none");
  const bool tc_path = use_tc && (edge_tc_features() & 2);
  if (dx_only && !tc_path) {   // the generic kernel always accumulates them: give it a sink
    float* sink = w.wg_scratch;
    dw0 = sink; sink += (size_t)H0 * a.ldwef;
    db0 = sink; sink += H0;
    dw1 = sink; sink += (size_t)H1 * H0;
    db1 = sink; sink += H1;
    dw2 = sink; sink += (size_t)H2 * H1;
    db2 = sink;
  }
  const size_t BN = (size_t)B * N;
  a.dagg = dagg;
  a.dW1 = dw1; a.db1 = db1; a.dW2 = dw2; a.db2 = db2;
  a.dP = w.dP; a.dQ = w.dQ; a.dx_ef = w.dxef;
  a.dWef = dw0 ? dw0 + 2 * F : nullptr;
  const bool tc_bwd = tc_path;
  if (tc_bwd) {   // dP | dQ adjacent: one clear
    MPG_CUDA(cudaMemsetAsync(w.dP, 0, (size_t)((char*)w.dQ - (char*)w.dP) + BN * H0 * sizeof(float), s));
  } else {
    MPG_CUDA(cudaMemsetAsync(w.dQ, 0, BN * H0 * sizeof(float), s));
  }
  if (a.n_ef) MPG_CUDA(cudaMemsetAsync(w.dxef, 0, BN * F * sizeof(float), s));
  if (tc_bwd) {
    if (launch_edge_tc_bwd(a, w.tc, w.tc_scratch, saved != nullptr, s)) return 1;
  } else {
    if (use_tc) {  // forward ran on tensor cores, backward kernel not built: generic needs W^T copies
      if (launch_transpose(w1, H1, H0, w.W1t, s)) return 1;
      if (launch_transpose(w2, H2, H1, w.W2t, s)) return 1;
    }
    if (launch_edge_generic(a, true, s)) return 1;
  }
  const bool precise = precision == 0;
  // node-level tail of the factorised first layer
  if (pq_supported(F, H0)) {
    if (a.cmap != nullptr)   // rows outside every tile (padded particles) get no gradient
      MPG_CUDA(cudaMemset2DAsync(dx, (size_t)lddx * sizeof(float), 0, (size_t)F * sizeof(float), BN, s));
    if (launch_pq_bwd(w.dP, w.dQ, x, ldx, w0, a.ldwef, dx, lddx, dw0, db0, (int)BN, F, H0, s, a.p_tiled != 0, tc_bwd,
                      tc_bwd ? a.cmap : nullptr, a.ctiles_max))
      return 1;
  } else {
    GemmEpi acc;
    acc.accumulate = 1;
    if (dw0 != nullptr) {   // null on the tcgen05 path of an input-gradient-only backward (frozen weights)
      if (launch_colsum(w.dP, H0, (int)BN, H0, db0, s)) return 1;
      int split = cdiv((long long)BN, 256);
      if (split > 64) split = 64;
      if (launch_gemm(false, false, precise, w.dP, H0, x, ldx, dw0, a.ldwef, H0, F, (int)BN, acc, split, s)) return 1;
      if (launch_gemm(false, false, precise, w.dQ, H0, x, ldx, dw0 + F, a.ldwef, H0, F, (int)BN, acc, split, s)) return 1;
    }
    GemmEpi e0;
    if (launch_gemm(true, false, precise, w.dP, H0, w0, a.ldwef, dx, lddx, (int)BN, F, H0, e0, 1, s)) return 1;
    if (launch_gemm(true, false, precise, w.dQ, H0, w0 + F, a.ldwef, dx, lddx, (int)BN, F, H0, acc, 1, s)) return 1;
  }
  if (a.n_ef) {
    add_strided_kernel<<<cdiv((long long)BN * F, 256), 256, 0, s>>>(dx, lddx, w.dxef, F, (int)BN, F);
    MPG_LAUNCH_CHECK();
  }
  return 0;
}

int mpg_edge_bwd(const float* x, int ldx, const float* mask, const float* w0, const float* b0, const float* w1,
                 const float* b1, const float* w2, const float* b2, int B, int N, int F, int H0, int H1, int H2,
                 int ef_mode, int nd, int mean, float alpha, float p_drop, uint64_t seed, const uint64_t* seed_dev,
                 int precision, void* workspace, size_t workspace_bytes, const float* dagg, float* dx, int lddx, float* dw0,
                 float* db0, float* dw1, float* db1, float* dw2, float* db2, void* stream) {
  return edge_bwd_impl(EdgeExtra(), nullptr, 0, x, ldx, mask, w0, b0, w1, b1, w2, b2, B, N, F, H0, H1, H2, ef_mode, nd,
                       mean, alpha, p_drop, seed, seed_dev, precision, workspace, workspace_bytes, dagg, dx, lddx, dw0, db0,
                       dw1, db1, dw2, db2, stream);
}

int mpg_edge_nbr_bwd(const int* nbr, int K, int knn_scale, const float* lc, const float* x, int ldx, const float* mask,
                     const float* w0, const float* b0, const float* w1, const float* b1, const float* w2, const float* b2,
                     int B, int N, int F, int H0, int H1, int H2, int ef_mode, int nd, int mean, float alpha, float p_drop,
                     uint64_t seed, const uint64_t* seed_dev, void* workspace, size_t workspace_bytes, const float* dagg,
                     float* dx, int lddx, float* dmask, float* dlc, float* dw0, float* db0, float* dw1, float* db1,
                     float* dw2, float* db2, void* stream) {
  EdgeExtra ex;
  ex.fp32_only = true;
  ex.nbr = nbr; ex.K = K; ex.knn_scale = knn_scale; ex.dmask = dmask; ex.lc = lc; ex.dlc = dlc;
  return edge_bwd_impl(ex, nullptr, 0, x, ldx, mask, w0, b0, w1, b1, w2, b2, B, N, F, H0, H1, H2, ef_mode, nd, mean, alpha,
                       p_drop, seed, seed_dev, 0, workspace, workspace_bytes, dagg, dx, lddx, dw0, db0, dw1, db1, dw2, db2,
                       stream);
}

int mpg_edge_bwd_saved(const void* fwd_workspace, size_t fwd_workspace_bytes, const float* x, int ldx,
                       const float* mask, const float* w0, const float* b0, const float* w1, const float* b1,
                       const float* w2, const float* b2, int B, int N, int F, int H0, int H1, int H2, int ef_mode,
                       int nd, int mean, float alpha, float p_drop, uint64_t seed, const uint64_t* seed_dev,
                       int precision, void* workspace, size_t workspace_bytes, const float* dagg, float* dx, int lddx,
                       float* dw0, float* db0, float* dw1, float* db1, float* dw2, float* db2, void* stream) {
  MPG_CHECK(fwd_workspace != nullptr, "edge_bwd_saved: null forward workspace");
  return edge_bwd_impl(EdgeExtra(), fwd_workspace, fwd_workspace_bytes, x, ldx, mask, w0, b0, w1, b1, w2, b2, B, N, F, H0, H1, H2,
                       ef_mode, nd, mean, alpha, p_drop, seed, seed_dev, precision, workspace, workspace_bytes, dagg, dx,
                       lddx, dw0, db0, dw1, db1, dw2, db2, stream);
}

int mpg_rank_mask(const float* x, int ldx, const float* labels, int ldl, int B, int N, float* mask, void* stream) {
  MPG_CHECK(N > 0 && N <= 8192, "rank_mask: N out of range");
  return launch_rank_mask(x, ldx, labels, ldl, B, N, mask, (cudaStream_t)stream);
}
int mpg_particle_order(const float* mask, int B, int N, int* pos, float* mask_sorted, void* stream) {
  return launch_particle_order(mask, B, N, pos, mask_sorted, (cudaStream_t)stream);
}
int mpg_batch_order(const float* key, int ldk, int B, int* pos, void* stream) {
  return launch_batch_order(key, ldk, B, pos, (cudaStream_t)stream);
}
int mpg_permute_rows(const float* src, int lds, float* dst, int ldd, const int* pos, int B, int N, int F, int mode,
                     void* stream) {
  MPG_CHECK(mode == 0 || mode == 1, "permute_rows: mode must be 0 (scatter) or 1 (gather)");
  return launch_permute_rows(src, lds, dst, ldd, pos, B, N, F, mode, (cudaStream_t)stream);
}
int mpg_ls_loss_fwd(const float* d, int n, int n0, float t0, float t1, float* loss, void* stream) {
  MPG_CHECK(n0 >= 1 && n0 <= n, "ls_loss: need 1 <= n0 <= n (n0 = %d, n = %d)", n0, n);
  return launch_ls_loss(d, nullptr, n, n0, t0, t1, loss, false, (cudaStream_t)stream);
}
int mpg_ls_loss_bwd(const float* d, const float* gout, int n, int n0, float t0, float t1, float* dd, void* stream) {
  MPG_CHECK(n0 >= 1 && n0 <= n, "ls_loss: need 1 <= n0 <= n (n0 = %d, n = %d)", n0, n);
  return launch_ls_loss(d, gout, n, n0, t0, t1, dd, true, (cudaStream_t)stream);
}
int mpg_split_mask(const float* x, int ldx, int rows, float* mask, void* stream) {
  return launch_split_mask(x, ldx, rows, mask, (cudaStream_t)stream);
}
int mpg_gen_tail_fwd(const float* h, const float* mask, float* out, int rows, int Fo, int act, void* stream) {
  return launch_gen_tail_fwd(h, mask, out, rows, Fo, act, (cudaStream_t)stream);
}
int mpg_gen_tail_bwd(const float* dout, const float* out, float* dh, int rows, int Fo, int ldo, int act,
                     void* stream) {
  return launch_gen_tail_bwd(dout, out, dh, rows, Fo, ldo, act, (cudaStream_t)stream);
}
int mpg_pool_fwd(const float* h, const float* mask, float* out, int B, int N, int C, int mean, void* stream) {
  return launch_pool_fwd(h, mask, out, B, N, C, mean, (cudaStream_t)stream);
}
int mpg_pool_bwd(const float* dout, const float* mask, float* dh, int B, int N, int C, int mean, void* stream) {
  return launch_pool_bwd(dout, mask, dh, B, N, C, mean, (cudaStream_t)stream);
}
int mpg_unary_fwd(const float* x, float* y, size_t n, int act, void* stream) {
  return launch_unary(x, nullptr, y, n, act, false, (cudaStream_t)stream);
}
int mpg_unary_bwd(const float* dy, const float* y, float* dx, size_t n, int act, void* stream) {
  return launch_unary(y, dy, dx, n, act, true, (cudaStream_t)stream);
}
int mpg_sn_fwd(const float* w_bar, float* u, float* v, float* w_out, float* sigma, int H, int W, void* stream) {
  MPG_CHECK((size_t)(H + W + 32) * 4 <= 48 * 1024, "spectral norm: weight too large");
  return launch_sn_fwd(w_bar, u, v, w_out, sigma, H, W, (cudaStream_t)stream);
}
int mpg_sn_bwd(const float* dw, const float* w_bar, const float* u, const float* v, const float* sigma,
               float* dw_bar, int H, int W, void* stream) {
  return launch_sn_bwd(dw, w_bar, u, v, sigma, dw_bar, H, W, (cudaStream_t)stream);
}
int mpg_rmsprop(float* p, const float* g, float* sq, size_t n, float lr, float alpha, float eps, float gscale,
                void* stream) {
  return launch_rmsprop(p, g, sq, n, lr, alpha, eps, gscale, (cudaStream_t)stream);
}

int mpg_attn_fwd(const float* q, int ldq, const float* k, int ldk, const float* v, int ldv, const float* key_mask,
                 int B, int Nq, int Nk, int E, int heads, float* o, float* p_saved, void* stream) {
  AttnArgs a{q, ldq, k, ldk, v, ldv, key_mask, B, Nq, Nk, E, heads};
  return launch_attn_fwd(a, o, p_saved, (cudaStream_t)stream);
}
int mpg_attn_bwd(const float* q, int ldq, const float* k, int ldk, const float* v, int ldv, const float* key_mask,
                 int B, int Nq, int Nk, int E, int heads, const float* p_saved, const float* dout, float* dq,
                 float* dk, float* dv, void* stream) {
  AttnArgs a{q, ldq, k, ldk, v, ldv, key_mask, B, Nq, Nk, E, heads};
  return launch_attn_bwd(a, p_saved, dout, dq, dk, dv, (cudaStream_t)stream);
}
int mpg_residual_dropout_fwd(const float* x, const float* r, float* out, size_t rows, int cols, float p_drop,
                             uint64_t seed, const uint64_t* seed_dev, uint32_t rng_stream, void* stream) {
  MPG_CHECK(p_drop >= 0.f && p_drop < 1.f, "dropout p must be in [0,1)");
  return launch_resdrop(x, r, out, rows, cols, make_drop(p_drop, seed, seed_dev), rng_stream, false,
                        (cudaStream_t)stream);
}
int mpg_residual_dropout_bwd(const float* dout, float* dx, size_t rows, int cols, float p_drop, uint64_t seed,
                             const uint64_t* seed_dev, uint32_t rng_stream, void* stream) {
  return launch_resdrop(dout, nullptr, dx, rows, cols, make_drop(p_drop, seed, seed_dev), rng_stream, true,
                        (cudaStream_t)stream);
}

// ---- second-order products of the fused edge op (double backward, WGAN-GP) ------------------------------------
size_t mpg_edge_bwd2_workspace_bytes(int B, int N, int F, int H0, int H1, int H2) {
  return carve_edge_ws(nullptr, B, N, F, H0, H1, H2).total + 2 * align_up((size_t)B * N * H0 * sizeof(float)) + align_up(H0 * sizeof(float));
}

int mpg_edge_bwd2(const float* x, int ldx, const float* u, int ldu, const float* mask, const float* w0, const float* b0,
                  const float* w1, const float* b1, const float* w2, const float* b2, int B, int N, int F, int H0, int H1,
                  int H2, int mean, float alpha, float p_drop, uint64_t seed, const uint64_t* seed_dev, void* workspace,
                  size_t workspace_bytes, const float* dagg, float* tagg, float* gmask, float* dw0, float* dw1, float* dw2,
                  void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  MPG_CHECK(workspace != nullptr && workspace_bytes >= mpg_edge_bwd2_workspace_bytes(B, N, F, H0, H1, H2),
            "edge_bwd2 workspace too small");
  MPG_CHECK(u && dagg && tagg, "edge_bwd2: null operand");
  EdgeArgs a;
  EdgeWs w;
  bool use_tc = false;
  // fp32 kernels: P/Q of the primal point, transposed weight copies
  if (edge_common(a, w, x, ldx, mask, w0, b0, w1, b1, w2, b2, B, N, F, H0, H1, H2, 0, 0, mean, alpha, p_drop, seed,
                  seed_dev, 0, workspace, workspace_bytes, &use_tc, s, false))
    return 1;
  const size_t BN = (size_t)B * N;
  char* extra = reinterpret_cast<char*>(workspace) + w.total;
  float* Pt = reinterpret_cast<float*>(extra);
  float* Qt = reinterpret_cast<float*>(extra + align_up(BN * H0 * sizeof(float)));
  float* zero_b = reinterpret_cast<float*>(extra + 2 * align_up(BN * H0 * sizeof(float)));
  // the first layer applied to the direction u (no bias: a constant has no tangent)
  MPG_CUDA(cudaMemsetAsync(zero_b, 0, H0 * sizeof(float), s));
  if (pq_supported(F, H0)) {
    if (launch_pq_fwd(u, ldu, w0, a.ldwef, zero_b, Pt, Qt, (int)BN, F, H0, s)) return 1;
  } else {
    GemmEpi e;
    if (launch_gemm(true, true, true, u, ldu, w0, a.ldwef, Pt, H0, (int)BN, H0, F, e, 1, s)) return 1;
    if (launch_gemm(true, true, true, u, ldu, w0 + F, a.ldwef, Qt, H0, (int)BN, H0, F, e, 1, s)) return 1;
  }
  a.dagg = dagg;
  a.Pt = Pt; a.Qt = Qt; a.tagg = tagg; a.gmask = gmask;
  // weight-gradient sinks when the caller wants only tagg / gmask
  float* sink = w.wg_scratch;
  float* sdw0 = sink; sink += (size_t)H0 * a.ldwef;
  float* sdb0 = sink; sink += H0;
  float* sdw1 = sink; sink += (size_t)H1 * H0 + H1;
  float* sdw2 = sink;
  a.dW1 = dw1 ? dw1 : sdw1;
  a.dW2 = dw2 ? dw2 : sdw2;
  a.db1 = nullptr; a.db2 = nullptr;   // biases have no second-order term
  a.dP = w.dP; a.dQ = w.dQ;
  MPG_CUDA(cudaMemsetAsync(w.dQ, 0, BN * H0 * sizeof(float), s));
  if (gmask != nullptr) MPG_CUDA(cudaMemsetAsync(gmask, 0, BN * sizeof(float), s));
  if (launch_edge_generic_tangent(a, s)) return 1;
  // dW0(2nd) = [dP^T u | dQ^T u] with the first-order dP / dQ; the dx and db0 the node-level kernel also produces go
  // to scratch
  if (dw0 != nullptr) {
    if (pq_supported(F, H0)) {
      if (launch_pq_bwd(w.dP, w.dQ, u, ldu, w0, a.ldwef, w.dxef, F, dw0, sdb0, (int)BN, F, H0, s)) return 1;
    } else {
      int split = cdiv((long long)BN, 256);
      if (split > 64) split = 64;
      GemmEpi acc;
      acc.accumulate = 1;
      if (launch_gemm(false, false, true, w.dP, H0, u, ldu, dw0, a.ldwef, H0, F, (int)BN, acc, split, s)) return 1;
      if (launch_gemm(false, false, true, w.dQ, H0, u, ldu, dw0 + F, a.ldwef, H0, F, (int)BN, acc, split, s)) return 1;
    }
  }
  (void)sdw0;
  return 0;
}

int mpg_knn_select(const float* x, int ldx, const float* mask, int B, int N, int nd, int k, int self_loops, int* idx,
                   void* stream) {
  MPG_CHECK(N > 0 && N <= 1024 && nd > 0, "knn_select: need 0 < N <= 1024 and nd > 0");
  const int skip = self_loops ? 0 : 1;
  MPG_CHECK(k > 0 && k + skip <= N, "knn_select: k + skipped self loop must not exceed N (k = %d, N = %d)", k, N);
  return launch_knn_select(x, ldx, mask, B, N, nd, k, skip, idx, (cudaStream_t)stream);
}

int mpg_gen_postprocess(const float* jets, int ldj, float* out, int ldo, size_t rows, int nfeat, const float* shift,
                        const float* norm, const float* maxv, int use_mask, void* stream) {
  MPG_CHECK(nfeat > 0 && nfeat <= 8 && nfeat < ldj + (use_mask ? 0 : 1) && nfeat <= ldo, "gen_postprocess: bad feature count");
  PostCfg c;
  memset(&c, 0, sizeof(c));
  c.nfeat = nfeat;
  for (int i = 0; i < nfeat; ++i) {   // host arrays; NaN = "None" in gen.py's lists (skip the step)
    if (shift != nullptr && shift[i] == shift[i] && shift[i] != 0.f) { c.shift[i] = shift[i]; c.has_shift |= 1u << i; }
    if (norm != nullptr && norm[i] == norm[i]) {
      MPG_CHECK(maxv != nullptr, "gen_postprocess: norm without maxes");
      c.norm[i] = norm[i]; c.maxv[i] = maxv[i]; c.has_norm |= 1u << i;
    }
  }
  return launch_gen_postprocess(jets, ldj, out, ldo, rows, c, use_mask, (cudaStream_t)stream);
}

int mpg_cond_columns(const float* x, int ldx, const float* cond, int C, float* out, size_t rows, int F, int B, void* stream) {
  MPG_CHECK(C > 0 && F > 0 && B > 0, "cond_columns: bad sizes");
  return launch_cond_columns(x, ldx, cond, C, out, rows, F, B, (cudaStream_t)stream);
}
size_t mpg_compact_map_ints(int B, int N) { return compact_map_ints(B, N); }
int mpg_compact_map(const float* mask, int B, int N, int* cmap, int* scratch, void* stream) {
  MPG_CHECK(mask != nullptr && cmap != nullptr && scratch != nullptr && B > 0 && N > 0, "compact_map: bad arguments");
  return launch_compact_map(mask, B, N, cmap, scratch, (cudaStream_t)stream);
}
int mpg_edge_set_compaction(const int* cmap) {
  g_cmap = cmap;
  return 0;
}
int mpg_split_mask_bwd(const float* dmask, float* dx, int ldx, size_t rows, void* stream) {
  return launch_split_mask_bwd(dmask, dx, ldx, rows, (cudaStream_t)stream);
}
int mpg_pool_dmask(const float* h, const float* dout, float* dmask, int B, int N, int C, float scale, void* stream) {
  return launch_pool_dmask(h, dout, dmask, B, N, C, scale, (cudaStream_t)stream);
}
int mpg_layernorm_fwd(const float* x, const float* w, const float* b, float* y, float* mean, float* rstd, size_t rows,
                      int C, float eps, void* stream) {
  MPG_CHECK(C > 0 && C <= 4096, "layernorm: C out of range");
  return launch_layernorm_fwd(x, w, b, y, mean, rstd, rows, C, eps, (cudaStream_t)stream);
}
int mpg_layernorm_bwd(const float* dy, const float* x, const float* w, const float* mean, const float* rstd, float* dx,
                      float* dw, float* db, size_t rows, int C, void* stream) {
  MPG_CHECK(C > 0 && C <= 4096, "layernorm: C out of range");
  MPG_CHECK((dw == nullptr) == (db == nullptr), "layernorm_bwd: pass both parameter gradients or none");
  return launch_layernorm_bwd(dy, x, w, mean, rstd, dx, dw, db, rows, C, (cudaStream_t)stream);
}

int mpg_mab_supported(int E, int heads, int Nq, int Nk) { return mab_supported(E, heads, Nq, Nk) ? 1 : 0; }
size_t mpg_mab_workspace_bytes(int B) { return mab_workspace_bytes(B); }

static int mab_args(MabArgs& a, const float* x, int ldx, const float* y, int ldy, const float* key_mask, const float* w_in,
                    const float* b_in, const float* w_out, const float* b_out, const float* w_ff, const float* b_ff, int B,
                    int Nq, int Nk, int E, int heads, float alpha, float p_res, float p_ff, uint64_t seed,
                    const uint64_t* seed_dev, float* q, float* kv, float* o, float* h, float* f, float* out) {
  MPG_CHECK(mab_supported(E, heads, Nq, Nk), "mab: unsupported shape E=%d heads=%d Nq=%d Nk=%d", E, heads, Nq, Nk);
  MPG_CHECK(p_res >= 0.f && p_res < 1.f && p_ff >= 0.f && p_ff < 1.f, "mab: dropout p must be in [0,1)");
  MPG_CHECK(x && y && w_in && b_in && w_out && b_out && w_ff && b_ff && q && kv && o && h && f, "mab: null operand");
  memset(&a, 0, sizeof(a));
  a.x = x; a.ldx = ldx; a.y = y; a.ldy = ldy; a.key_mask = key_mask;
  a.w_in = w_in; a.b_in = b_in; a.w_out = w_out; a.b_out = b_out; a.w_ff = w_ff; a.b_ff = b_ff;
  a.B = B; a.Nq = Nq; a.Nk = Nk; a.alpha = alpha;
  a.drop_res = make_drop(p_res, seed, seed_dev);
  a.drop_ff = make_drop(p_ff, seed, seed_dev);
  a.q = q; a.kv = kv; a.o = o; a.h = h; a.f = f; a.out = out;
  return 0;
}

int mpg_mab_fwd(const float* x, int ldx, const float* y, int ldy, const float* key_mask, const float* w_in,
                const float* b_in, const float* w_out, const float* b_out, const float* w_ff, const float* b_ff, int B,
                int Nq, int Nk, int E, int heads, float alpha, float p_res, float p_ff, uint64_t seed,
                const uint64_t* seed_dev, int precision, void* workspace, size_t workspace_bytes, float* q, float* kv,
                float* o, float* h, float* f, float* out, void* stream) {
  MabArgs a;
  if (mab_args(a, x, ldx, y, ldy, key_mask, w_in, b_in, w_out, b_out, w_ff, b_ff, B, Nq, Nk, E, heads, alpha, p_res, p_ff,
               seed, seed_dev, q, kv, o, h, f, out))
    return 1;
  MPG_CHECK(out != nullptr && workspace != nullptr && workspace_bytes >= mab_workspace_bytes(B), "mab_fwd: workspace too small");
  return launch_mab_fwd(a, workspace, precision, (cudaStream_t)stream);
}

int mpg_mab_bwd(const float* x, int ldx, const float* y, int ldy, const float* key_mask, const float* w_in,
                const float* b_in, const float* w_out, const float* b_out, const float* w_ff, const float* b_ff, int B,
                int Nq, int Nk, int E, int heads, float alpha, float p_res, float p_ff, uint64_t seed,
                const uint64_t* seed_dev, int precision, void* workspace, size_t workspace_bytes, const float* q,
                const float* kv, const float* o, const float* h, const float* f, const float* dout, float* dx, float* dy,
                float* dw_in, float* db_in, float* dw_out, float* db_out, float* dw_ff, float* db_ff, void* stream) {
  MabArgs a;
  if (mab_args(a, x, ldx, y, ldy, key_mask, w_in, b_in, w_out, b_out, w_ff, b_ff, B, Nq, Nk, E, heads, alpha, p_res, p_ff,
               seed, seed_dev, const_cast<float*>(q), const_cast<float*>(kv), const_cast<float*>(o), const_cast<float*>(h),
               const_cast<float*>(f), nullptr))
    return 1;
  MPG_CHECK(workspace != nullptr && workspace_bytes >= mab_workspace_bytes(B), "mab_bwd: workspace too small");
  const bool self = x == y && ldx == ldy && Nq == Nk;
  MPG_CHECK(dout && dx && (self || dy), "mab_bwd: null gradient pointer");
  const bool none = !dw_in && !db_in && !dw_out && !db_out && !dw_ff && !db_ff;
  MPG_CHECK(none || (dw_in && db_in && dw_out && db_out && dw_ff && db_ff), "mab_bwd: pass all six parameter gradients or none");
  MabGrads g;
  memset(&g, 0, sizeof(g));
  g.dout = dout; g.dx = dx; g.dy = dy;
  g.dw_in = dw_in; g.db_in = db_in; g.dw_out = dw_out; g.db_out = db_out; g.dw_ff = dw_ff; g.db_ff = db_ff;
  return launch_mab_bwd(a, g, workspace, precision, (cudaStream_t)stream);
}

size_t mpg_peer_flag_words(int ctas, int world) { return peer_flag_words(ctas, world); }

int mpg_allreduce_rmsprop(float* p, float* sq, const void* const* peer_grads, void* const* peer_flags,
                          const void* grads_multicast, size_t n, int rank, int world, int ctas, float lr, float alpha,
                          float eps, void* stream) {
  MPG_CHECK(peer_grads != nullptr && peer_flags != nullptr, "allreduce_rmsprop: null pointer table");
  MPG_CHECK(world >= 2 && world <= MPG_PEER_MAX, "allreduce_rmsprop: world size must be in [2, %d]", MPG_PEER_MAX);
  PeerArgs a;
  memset(&a, 0, sizeof(a));
  a.p = p; a.sq = sq; a.n = n; a.rank = rank; a.world = world; a.lr = lr; a.alpha = alpha; a.eps = eps;
  a.grads_mc = reinterpret_cast<const float*>(grads_multicast);
  MPG_CHECK((reinterpret_cast<uintptr_t>(grads_multicast) & 15) == 0, "allreduce_rmsprop: unaligned multicast pointer");
  for (int r = 0; r < world; ++r) {
    a.grads[r] = reinterpret_cast<const float*>(peer_grads[r]);
    a.flags[r] = reinterpret_cast<unsigned*>(peer_flags[r]);
  }
  return launch_allreduce_rmsprop(a, ctas, (cudaStream_t)stream);
}

}  // extern "C"
