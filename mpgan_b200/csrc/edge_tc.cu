// tcgen05 / TMEM edge network for the default MPGAN architecture (fe = 96 -> 160 -> 192): host-side
// launchers.  Kernels: edge_tc_fwd.cuh (forward), edge_tc_bwd.cuh (backward by recompute); shared PTX
// wrappers, tile layouts and the weight-image kernel: edge_tc_common.cuh.
#include <cstdlib>

#include "edge_node.cuh"

namespace mpg {
namespace {

#include "edge_tc_common.cuh"
#include "edge_tc_fwd.cuh"
#include "edge_tc_bwd.cuh"

}  // namespace

int edge_tc_features() { return 3; }

#ifdef MPG_TRACE
extern "C" int mpg_debug_set_trace(void* p) {
  return (int)cudaMemcpyToSymbol(g_trace, &p, sizeof(p));
}
#endif

// one-shot timing probes: bench.py arms a pair of CUDA events per kernel id (1 = forward,
// 2 = backward CHAIN, 3 = backward DW2); the next launch of that kernel is bracketed by them
static thread_local cudaEvent_t g_probe[4][2] = {};
void edge_tc_arm_probe(int id, cudaEvent_t e0, cudaEvent_t e1) {
  if (id >= 1 && id <= 3) { g_probe[id][0] = e0; g_probe[id][1] = e1; }
}
struct ProbeScope {
  int id; cudaStream_t s; bool on;
  ProbeScope(int id_, cudaStream_t s_) : id(id_), s(s_), on(g_probe[id_][0] != nullptr) {
    if (on) cudaEventRecord(g_probe[id][0], s);
  }
  ~ProbeScope() {
    if (on) { cudaEventRecord(g_probe[id][1], s); g_probe[id][0] = g_probe[id][1] = nullptr; }
  }
};

bool edge_tc_supported(const EdgeArgs& a) {
  // a 128-row tile may span at most F_QJ jets (Q ring of the forward kernel)
  return a.H0 == K0 && a.H1 == N1 && a.H2 == N2 && a.n_ef == 0 && (TILE - 1) / a.N + 2 <= F_QJ &&
         (a.drop.p == 0.f || a.drop.p == 0.5f) && a.alpha > 0.f && a.alpha < 1.f;
}

static size_t tc_sbits_bytes(int B, int N) {   // backward: one uint2 per (step, epilogue thread)
  const long long tiles = compact_tiles_max(B, N);   // (covers the uncompacted ceil(B*N / 128) too)
  return (size_t)tiles * N * F_NEPI * sizeof(uint2);
}

static int tc_num_sms() {
  int dev = 0, sms = 148;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  return sms;
}
static size_t tc_slab_bytes() { return (size_t)tc_num_sms() * SLAB_FLOATS * sizeof(float); }

static size_t tc_steps_bytes(int B, int N) {   // work list: int2 per (tile, sender) + the total
  const long long tiles = compact_tiles_max(B, N);
  return ((size_t)tiles * N * sizeof(int2) + 256 + 255) & ~(size_t)255;
}

size_t edge_tc_persist_bytes(int B, int N, int H0, int H1, int H2) {
  if (H0 != K0 || H1 != N1 || H2 != N2) return 0;
  return W1_BYTES + W2_BYTES + 1024 + tc_steps_bytes(B, N);
}
size_t edge_tc_scratch_bytes(int B, int N, int H0, int H1, int H2) {
  if (H0 != K0 || H1 != N1 || H2 != N2) return 0;
  return tc_sbits_bytes(B, N) + 512 + tc_slab_bytes();
}

// the activation / gradient tiles hold X / (sd * sl) (dropout and leaky-relu scales), the weight images sd * sl * W
static int tc_prepare(const EdgeArgs& a, void* persist, void* scratch, bool reuse, TcArgs& t, int* grid,
                      cudaStream_t stream, float* zero_agg = nullptr) {
  MPG_CHECK(edge_tc_supported(a), "edge_tc: unsupported configuration");
  uint8_t* img = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(persist) + 255) & ~(uintptr_t)255);
  t.a = a;
  t.w1img = img;
  t.w2img = img + W1_BYTES;
  uint8_t* p = img + W1_BYTES + W2_BYTES;
  int* total = reinterpret_cast<int*>(p);
  t.total_steps = total;
  t.steps = reinterpret_cast<int2*>(p + 256);
  t.sbits = nullptr;
  t.wslab = nullptr;
  if (scratch != nullptr) {
    uint8_t* q = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(scratch) + 255) & ~(uintptr_t)255);
    t.sbits = reinterpret_cast<uint2*>(q);
    t.wslab = reinterpret_cast<float*>(q + ((tc_sbits_bytes(a.B, a.N) + 255) & ~(size_t)255));
  }
  const long long BN = (long long)a.B * a.N;
  t.num_tiles = a.cmap ? a.ctiles_max : (int)((BN + TILE - 1) / TILE);   // compaction: upper bound (actual count on the device)
  if (!reuse) {   // images and work list of this (weights, mask): the backward of the same call finds them here
    const float s = (a.drop.p > 0.f ? 2.f : 1.f) * 0.5f * (1.f + a.alpha);
    const size_t agg_floats = zero_agg != nullptr ? (size_t)a.B * a.N * N2 : 0;   // N2 % 4 == 0; torch buffers are 16-byte aligned
    MPG_CHECK((reinterpret_cast<uintptr_t>(zero_agg) & 15) == 0, "edge_tc: agg must be 16-byte aligned");
    PqFwdArgs pq{};
    int nb_pq = 0;
    const long long BNl = (long long)a.B * a.N;
    // The set-up kernel's last block builds the work list when that is quick (a warp per tile and 32 senders: ~0.15 us
    // each on one SM); larger problems use the many-block kernel after it.  MPG_STEP_LIST_MAX_UNITS: test knob.
    long long list_max = 384;
    if (const char* e = getenv("MPG_STEP_LIST_MAX_UNITS")) list_max = atoll(e);
    const bool list_in_block = (long long)t.num_tiles * ((a.N + 31) / 32) <= list_max &&
                               step_list_smem(BNl, a.N) <= 160 * 1024;
    size_t smem = list_in_block ? step_list_smem(BNl, a.N) : 0;
    if (a.pq_deferred) {   // P / Q of this call: first layer's node-level ends (edge_node.cuh)
      pq = PqFwdArgs{a.x, a.ldx, a.Wef - 2 * a.F, a.ldwef, a.b0, const_cast<float*>(a.P), const_cast<float*>(a.Q),
                     (int)BNl, a.F, a.H0, a.p_tiled, a.cmap, a.ctiles_max};
      nb_pq = 2 * (a.cmap ? a.ctiles_max : cdiv(BNl, PQ_ROWS));   // a block per (128-row tile, P | Q)
      if (pq_fwd_smem(a.F, a.H0) > smem) smem = pq_fwd_smem(a.F, a.H0);
    }
    const int nb_prep = cdiv(N1 * 128 + N2 * 192, 256);
    PrepArgs pr{a.W1, a.b1, a.W2, a.b2, s, img, img + W1_BYTES, reinterpret_cast<float4*>(zero_agg), agg_floats / 4};
    int stride = (int)(t.num_tiles * 0.381966f) | 1;   // ~ golden-ratio step: consecutive slots are far apart
    auto gcd = [](int x, int y) { while (y) { const int r = x % y; x = y; y = r; } return x; };
    while (stride > 1 && gcd(stride, t.num_tiles) != 1) stride -= 2;
    if (stride < 1) stride = 1;
    ListArgs ls{a.mask, a.B, a.N, t.num_tiles, const_cast<int2*>(t.steps), total, list_in_block ? 1 : 0, stride, a.cmap, a.ctiles_max};
    MPG_CUDA(cudaFuncSetAttribute(edge_setup_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    edge_setup_kernel<<<nb_pq + nb_prep + (list_in_block ? 1 : 0), 256, smem, stream>>>(pq, nb_pq, pr, nb_prep, ls);
    MPG_LAUNCH_CHECK();
    if (!list_in_block) {
      step_list_kernel<<<cdiv(t.num_tiles, 8), 256, 0, stream>>>(a.mask, a.B, a.N, t.num_tiles, const_cast<int2*>(t.steps), total,
                                                                 a.cmap, a.ctiles_max);
      MPG_LAUNCH_CHECK();
    }
  }
  // the number of live steps is only known on the device: one CTA per SM, CTAs without steps exit at once
  const long long max_steps = (long long)t.num_tiles * a.N;
  const int sms = tc_num_sms();
  *grid = (int)(max_steps < sms ? max_steps : sms);
  return 0;
}

int launch_edge_tc_fwd(const EdgeArgs& a, void* persist, cudaStream_t stream) {
  TcArgs t;
  int grid = 1;
  if (tc_prepare(a, persist, nullptr, false, t, &grid, stream, a.agg)) return 1;   // also zero-fills agg
  if (a.drop.p > 0.f) {
    MPG_CUDA(cudaFuncSetAttribute(edge_tc_fwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, F_SMEM));
    ProbeScope probe(1, stream);
    edge_tc_fwd_kernel<true><<<grid, F_NTHR, F_SMEM, stream>>>(t);
  } else {
    MPG_CUDA(cudaFuncSetAttribute(edge_tc_fwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, F_SMEM));
    ProbeScope probe(1, stream);
    edge_tc_fwd_kernel<false><<<grid, F_NTHR, F_SMEM, stream>>>(t);
  }
  MPG_LAUNCH_CHECK();
  return 0;
}

template <int MODE, bool DROP>
static int launch_bwd_one(const TcArgs& t, int grid, uint32_t smem, cudaStream_t stream) {
  MPG_CUDA(cudaFuncSetAttribute(edge_tc_bwd_kernel<MODE, DROP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  {
    ProbeScope probe(MODE == BWD_CHAIN ? 2 : 3, stream);
    edge_tc_bwd_kernel<MODE, DROP><<<grid, F_NTHR, smem, stream>>>(t);
  }
  MPG_LAUNCH_CHECK();
  return 0;
}

// dP and dQ must be zeroed by the caller (both are accumulated with atomics here)
int launch_edge_tc_bwd(const EdgeArgs& a, void* persist, void* scratch, bool reuse, cudaStream_t stream) {
  TcArgs t;
  int grid = 1;
  if (tc_prepare(a, persist, scratch, reuse, t, &grid, stream)) return 1;
  const bool wgrad = a.dW1 != nullptr;   // null: input gradient only -> CHAIN alone (its dW1 slab is simply not reduced)
  if (a.drop.p > 0.f) {
    if (launch_bwd_one<BWD_CHAIN, true>(t, grid, C_SMEM, stream)) return 1;
    if (wgrad && launch_bwd_one<BWD_DW2, true>(t, grid, D_SMEM, stream)) return 1;
  } else {
    if (launch_bwd_one<BWD_CHAIN, false>(t, grid, C_SMEM, stream)) return 1;
    if (wgrad && launch_bwd_one<BWD_DW2, false>(t, grid, D_SMEM, stream)) return 1;
  }
  if (wgrad) {
    wgrad_reduce_kernel<<<cdiv(SLAB_FLOATS / 4, WR_COLS), WR_COLS * WR_PARTS, 0, stream>>>(t.wslab, grid, a.dW1, a.db1, a.dW2, a.db2);
    MPG_LAUNCH_CHECK();
  }
  return 0;
}

}  // namespace mpg
