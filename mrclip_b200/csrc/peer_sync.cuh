// Device-side plumbing of the multi-GPU step over NVLink peer memory: flags instead of collectives.
//
// Every workspace owns a small symmetric (peer-mapped) control block `sig`: int32 [kPeerChannels][kPeerMaxRanks].
// Rank r announces "my data for channel c of step e has landed in your memory" by storing its epoch e into
// sig[c][r] of EVERY rank (release at system scope, after its peer stores); a consumer kernel spins (acquire at
// system scope) until the entries of the ranks it depends on have reached its own epoch of that channel.  Epochs live
// in ordinary device memory (`ctl`), are advanced by the producing kernel itself and read by the consuming kernels that
// follow it in the same stream -- nothing is baked into a launch, so a captured CUDA graph of the step replays correctly.
//
// The waiters spin on flags written by kernels running on OTHER GPUs (never on another kernel of the same GPU), so no
// co-scheduling assumption is made; a waiter that is not served within MRCLIP_SPIN_LIMIT_CYCLES traps instead of
// hanging the device.
//
// Channels: 0 = packed text rows landed (all-gather of loss.py:51-57), 1 = LSE statistics landed, 2 = text-gradient
// tiles landed (reduce-scatter of torch/distributed/nn/functional.py:343-347), 3/4 = scalars (loss / d logit_scale of the
// global-loss modes).
#pragma once
#include "ptx.cuh"

namespace mrclip {

constexpr int kPeerChannels = 8;
constexpr int kPeerMaxRanks = 64;
enum : int { CH_TEXT = 0, CH_STATS = 1, CH_DTEXT = 2, CH_LOSS = 3, CH_DSCALE = 4 };
// Layout of one rank's symmetric control block (kPeerBlockBytes, zero-initialised once):
//   int32 sig[kPeerChannels][kPeerMaxRanks] | float scal[2][kPeerMaxRanks] (loss, d logit_scale of every rank) |
//   float r2_in[kPeerMaxRanks] (entropy / R2 shares of my columns, one per source rank)
constexpr int kPeerScalOff = kPeerChannels * kPeerMaxRanks * 4;
constexpr int kPeerR2Off = kPeerScalOff + 2 * kPeerMaxRanks * 4;
constexpr int kPeerBlockBytes = 4096;

// ctl block (plain device memory, int32): epoch[kPeerChannels], then done[kPeerChannels] (block counters)
struct PeerCtl {
  int epoch[kPeerChannels];
  int done[kPeerChannels];
};

struct PeerInfo {
  const unsigned long long* sig_peers;   // [ranks] mapped addresses of every rank's sig block (own included)
  int* sig_local;                        // this rank's sig block
  PeerCtl* ctl;
  int ranks, rank;
};

__device__ __forceinline__ void st_release_sys(int* p, int v) {
  asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ int ld_acquire_sys(const int* p) {
  int v;
  asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ int ld_relaxed_gpu(const int* p) {
  int v;
  asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// generic-proxy acquire -> later async-proxy (TMA) reads of the data the flag guards
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }

__device__ __forceinline__ float* peer_scal(const PeerInfo& pi, int k) {     // rank k's scal block (mapped)
  return reinterpret_cast<float*>(__ldg(pi.sig_peers + k) + kPeerScalOff);
}
__device__ __forceinline__ float* peer_r2in(const PeerInfo& pi, int k) {     // rank k's r2_in block (mapped)
  return reinterpret_cast<float*>(__ldg(pi.sig_peers + k) + kPeerR2Off);
}
__device__ __forceinline__ const float* local_scal(const PeerInfo& pi) {
  return reinterpret_cast<const float*>(reinterpret_cast<const char*>(pi.sig_local) + kPeerScalOff);
}
__device__ __forceinline__ float* local_r2in(const PeerInfo& pi) {
  return reinterpret_cast<float*>(reinterpret_cast<char*>(pi.sig_local) + kPeerR2Off);
}

// Spin until sig_local[channel][src] >= want.  One thread.
__device__ __forceinline__ void peer_wait_one(const int* sig_local, int channel, int src, int want) {
  const int* f = sig_local + channel * kPeerMaxRanks + src;
  if (ld_acquire_sys(f) >= want) return;
  const long long t0 = clock64();
  while (ld_acquire_sys(f) < want) {
    if (clock64() - t0 > MRCLIP_SPIN_LIMIT_CYCLES) {
      printf("mrclip: peer flag wait timed out (channel %d, source rank %d, want epoch %d, have %d)\n", channel, src, want,
             ld_acquire_sys(f));
      __trap();
    }
  }
}

// Block-wide: wait until every rank's flag of `channel` has reached this rank's current epoch of that channel.
// Call before reading anything the peers pushed; all threads must call it.
__device__ __forceinline__ void peer_wait_all(const PeerInfo& pi, int channel) {
  if (pi.ranks > 1) {
    if ((int)threadIdx.x < pi.ranks) peer_wait_one(pi.sig_local, channel, threadIdx.x, ld_relaxed_gpu(&pi.ctl->epoch[channel]));
    __syncthreads();
  }
}

// Grid-wide "my pushes are done" -> flag on every rank.  Every thread of every block calls it after its last peer
// store; the last block to arrive advances the epoch and signals.  Blocks may exit right after.
__device__ __forceinline__ void peer_signal_when_grid_done(const PeerInfo& pi, int channel, int total_blocks) {
  __threadfence_system();            // this thread's peer stores are performed before the block's arrival below
  __syncthreads();
  if (threadIdx.x == 0) {
    const int prev = atomicAdd(&pi.ctl->done[channel], 1);
    if (prev == total_blocks - 1) {
      pi.ctl->done[channel] = 0;
      const int e = pi.ctl->epoch[channel] + 1;
      pi.ctl->epoch[channel] = e;
      __threadfence_system();        // (cumulativity: the other blocks' fenced stores, observed through the counter)
      for (int k = 0; k < pi.ranks; ++k)
        st_release_sys(reinterpret_cast<int*>(__ldg(pi.sig_peers + k)) + channel * kPeerMaxRanks + pi.rank, e);
    }
  }
}

}  // namespace mrclip
