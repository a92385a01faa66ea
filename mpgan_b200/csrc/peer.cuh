// Fused gradient all-reduce + RMSprop over NVLink peer memory (peer.cu).
#pragma once
#include "common.cuh"
namespace mpg {
constexpr int MPG_PEER_MAX = 16;
struct PeerArgs {
  float* p;                       // local flat parameters
  float* sq;                      // local RMSprop accumulators
  const float* grads[MPG_PEER_MAX];   // every rank's flat gradient buffer (symmetric memory), as mapped on this GPU
  unsigned* flags[MPG_PEER_MAX];      // every rank's flag block (symmetric memory, zero-initialised)
  const float* grads_mc;              // optional NVLink-SHARP multicast address of the gradient buffers: one
                                      // multimem.ld_reduce returns the sum over all ranks, reduced in the switch
  size_t n;
  int rank, world;
  float lr, alpha, eps;
};
size_t peer_flag_words(int ctas, int world);
int launch_allreduce_rmsprop(const PeerArgs& a, int ctas, cudaStream_t s);
}  // namespace mpg
