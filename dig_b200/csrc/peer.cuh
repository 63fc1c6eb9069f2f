// dig_b200 -- device-side pieces of the NVLink peer-memory collectives (see peer.cu for the protocol and the workspace layout).
#pragma once
#include <stdint.h>

#include "common.cuh"

namespace dig {

static constexpr int kPeerMaxRanks = 8;
static constexpr int kPeerChannels = 5;   // 0-2 SyncBatchNorm statistics, 3 keys, 4 gradient all-reduce
static constexpr int kPeerMaxFloats = 8192;                       // 2 x 4096 columns (projector width, M:463-482)
static constexpr long long kPeerWaitNs = 20ll * 1000 * 1000 * 1000;  // a peer that has not arrived after 20 s is gone: flag the error, do not hang

struct PeerTable {
  unsigned char* base[kPeerMaxRanks];
};

__host__ __device__ inline size_t peer_msg_off(int channel, int slot, int rank) {
  return ((size_t)(channel * 2 + slot) * kPeerMaxRanks + rank) * kPeerMaxFloats * sizeof(float);
}
__host__ __device__ inline size_t peer_flags_off() { return (size_t)kPeerChannels * 2 * kPeerMaxRanks * kPeerMaxFloats * sizeof(float); }
__host__ __device__ inline size_t peer_flag_off(int channel, int slot, int rank) {
  return peer_flags_off() + ((size_t)(channel * 2 + slot) * kPeerMaxRanks + rank) * sizeof(uint32_t);
}
__host__ __device__ inline size_t peer_ticket_off(int channel) {
  return peer_flags_off() + (size_t)kPeerChannels * 2 * kPeerMaxRanks * sizeof(uint32_t) + channel * sizeof(uint32_t);
}
__host__ __device__ inline size_t peer_err_off() { return peer_ticket_off(kPeerChannels); }
__host__ __device__ inline size_t peer_keys_off() { return (peer_err_off() + sizeof(uint32_t) + 1023) / 1024 * 1024; }

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) { asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float ld_relaxed_sys_f32(const float* p) {   // written by a peer GPU: never served from a stale L1 line
  float v;
  asm volatile("ld.relaxed.sys.global.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ long long global_ns() {
  long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// Whole thread block: raise flag[channel][slot][rank] = epoch on every peer (after this block's -- and, through the caller's ticket, the
// grid's -- data stores), then wait until every source rank's flag in OUR memory has reached epoch.
__device__ __forceinline__ void peer_signal_and_wait(const PeerTable& pt, int world, int rank, int channel, int slot, uint32_t epoch) {
  __threadfence_system();
  __syncthreads();
  if ((int)threadIdx.x < world) {
    st_release_sys(reinterpret_cast<uint32_t*>(pt.base[threadIdx.x] + peer_flag_off(channel, slot, rank)), epoch);
    const uint32_t* mine = reinterpret_cast<const uint32_t*>(pt.base[rank] + peer_flag_off(channel, slot, (int)threadIdx.x));
    const long long t0 = global_ns();
    while ((int32_t)(ld_acquire_sys(mine) - epoch) < 0) {
      if (global_ns() - t0 > kPeerWaitNs) {
        atomicExch(reinterpret_cast<unsigned int*>(pt.base[rank] + peer_err_off()), 1u);
        break;
      }
    }
  }
  __syncthreads();
}


// Whole thread block: buf[0..n) <- sum over ranks of buf, in rank order (bit-identical on every rank).  Optionally copies this rank's own
// values to loc0[0..nloc) / loc1[0..nloc) = buf[0..nloc) / buf[nloc..2 nloc) first (per-rank parameter gradients of BatchNorm).
struct PeerReduce {
  PeerTable pt;
  int world, rank, channel;   // world <= 1: no exchange
  uint32_t epoch;
  float* buf;
  int n;
  float* loc0;
  float* loc1;
  int nloc;
};

__device__ __forceinline__ void peer_allreduce_block(const PeerReduce& pr) {
  const int slot = (int)(pr.epoch & 1u);
  for (int i = threadIdx.x; i < pr.n; i += blockDim.x) {
    const float v = __ldcg(pr.buf + i);      // accumulated with L2 atomics by other blocks of this grid
    if (pr.loc0 != nullptr && i < pr.nloc) pr.loc0[i] = v;
    if (pr.loc1 != nullptr && i >= pr.nloc && i < 2 * pr.nloc) pr.loc1[i - pr.nloc] = v;
    for (int p = 0; p < pr.world; ++p) reinterpret_cast<float*>(pr.pt.base[p] + peer_msg_off(pr.channel, slot, pr.rank))[i] = v;
  }
  peer_signal_and_wait(pr.pt, pr.world, pr.rank, pr.channel, slot, pr.epoch);
  for (int i = threadIdx.x; i < pr.n; i += blockDim.x) {
    float s = 0.f;
    for (int r = 0; r < pr.world; ++r)
      s += ld_relaxed_sys_f32(reinterpret_cast<const float*>(pr.pt.base[pr.rank] + peer_msg_off(pr.channel, slot, r)) + i);
    pr.buf[i] = s;
  }
}

// Call at the very end of a multi-block kernel whose blocks accumulated into pr.buf with atomics: the last block to arrive runs the
// exchange (fused "column statistics + all-reduce": the collective costs no extra launch).
__device__ __forceinline__ void peer_allreduce_grid_tail(const PeerReduce& pr) {
  if (pr.world <= 1) return;
  __shared__ int is_last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned int* ticket = reinterpret_cast<unsigned int*>(pr.pt.base[pr.rank] + peer_ticket_off(pr.channel));
    const unsigned int total = gridDim.x * gridDim.y * gridDim.z;
    const unsigned int t = atomicAdd(ticket, 1u);
    is_last = (t == total - 1);
    if (is_last) *ticket = 0u;
  }
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  peer_allreduce_block(pr);
}

int fill_peer_table(PeerTable* pt, const int64_t* bases, int world);

}  // namespace dig
