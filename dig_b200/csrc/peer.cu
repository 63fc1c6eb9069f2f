// dig_b200 -- collectives of the data-parallel step carried by the kernels themselves over NVLink peer memory (no NCCL call):
//   * SyncBatchNorm statistics (R:390 converts all 14 BatchNorm1d; C3/C4 of SURVEY.md 2.3): a one-shot all-reduce of <= 2 x 4096 floats
//     per layer.  It is pure latency -- 22 of them sit on the critical path of the head chains every step -- so each rank PUSHES its partial
//     sums straight into every peer's inbox (st.global on the peer's mapped address, NVSwitch gives every peer full bandwidth), raises
//     a flag there, waits for the W flags in its OWN memory and adds the W partials in rank order (every rank gets the bit-identical
//     sum, and the sum does not depend on arrival order).  One small kernel, ~2 NVLink latencies, instead of an NCCL launch per layer.
//   * MoCo key all-gather (concat_all_gather, M:580-591; C1): the L2-normalised keys of this rank are written by the normalising kernel's
//     grid directly into every peer's key table, already in the [k1 of all ranks ; k2 of all ranks] order the logits GEMM consumes.
//
//   * Gradient averaging (DistributedDataParallel's bucket all-reduce, R:391; C2): the flat fp32 gradient buffer of every rank lives in an
//     IPC-mapped allocation.  One kernel per step: rank r owns the r-th slice of the buffer, waits until every rank's backward has
//     finished (flags), reads that slice from all W buffers over NVLink (128-bit loads), adds them in rank order, scales by 1/W and
//     stores the result into all W buffers; a second flag round makes kernel completion mean "my whole buffer is averaged".  The sum is
//     formed once per element, so every rank holds bit-identical gradients.
//
// Memory: every rank cudaMalloc's one workspace (dig_peer_alloc), the ranks exchange its IPC handle through torch.distributed and map
// each other's (dig_peer_open).  Layout of a workspace (floats unless noted), W = world size, kPeerMaxFloats per message:
//   [channel c][slot s in 0..1][source rank r] message area      (c < kPeerChannels)
//   flags  uint32 [channel][slot][source rank]                   epoch number of the message in that area
//   ticket uint32 [channel]                                      last-CTA election of the multi-CTA kernels
//   err    uint32                                                set when a wait timed out (a peer never arrived)
//   key table (bytes given by the caller) [slot][...]
// Slots alternate with the per-channel epoch: a rank can be at most one exchange ahead of a peer (it cannot finish exchange e+1
// before the peer has pushed e+1, which the peer does only after it has consumed e), so two slots are enough.  A channel is used by ONE
// stream per rank and in the same order on every rank (online forward / momentum forward / backward / keys).
#include <stdint.h>
#include <string.h>

#include "peer.cuh"
#include "../../include/dig_b200.h"

namespace dig {

// buf[0..n) <- sum over ranks of buf (in rank order).  One thread block.
__global__ void __launch_bounds__(256)
peer_allreduce_kernel(PeerReduce pr) { peer_allreduce_block(pr); }

// L2-normalise the rows of x [2Q, C] (this rank's [k1 ; k2], M:446-447 F.normalize) and write them into EVERY rank's key table at
// [half][rank * Q + i][C] (half = row / Q): the table then holds concat_all_gather(k1) followed by concat_all_gather(k2) (M:551-552,
// M:580-591).  Warp per row; the grid's last block raises the flags and waits for the other ranks, so that the kernel's completion means
// "all keys of all ranks are here".  local_copy (optional) receives this rank's normalised rows as well ([2Q, C]).
__global__ void __launch_bounds__(256)
peer_l2norm_allgather_kernel(PeerTable pt, const float* __restrict__ x, float* __restrict__ local_copy, long long Q, int C, int world, int rank,
                             int channel, uint32_t epoch, size_t table_bytes_per_slot) {
  const int slot = (int)(epoch & 1u);
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row < 2 * Q) {
    float s = 0.f;
    for (int c = lane; c < C; c += 32) { const float v = x[row * C + c]; s += v * v; }
    s = warp_sum(s);
    const float inv = 1.0f / fmaxf(sqrtf(s), 1e-12f);
    const long long half = row / Q, i = row % Q;
    const size_t off = peer_keys_off() + (size_t)slot * table_bytes_per_slot + ((half * world + rank) * Q + i) * (size_t)C * sizeof(float);
    for (int c = lane; c < C; c += 32) {
      const float v = x[row * C + c] * inv;
      if (local_copy) local_copy[row * C + c] = v;
      for (int p = 0; p < world; ++p) reinterpret_cast<float*>(pt.base[p] + off)[c] = v;
    }
  }
  // last block of the grid: every block's stores are ordered before its ticket (fence + atomic), so the flag follows all of them
  __shared__ int is_last;
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned int* ticket = reinterpret_cast<unsigned int*>(pt.base[rank] + peer_ticket_off(channel));
    const unsigned int t = atomicAdd(ticket, 1u);
    is_last = (t == gridDim.x - 1);
    if (is_last) *ticket = 0u;
  }
  __syncthreads();
  if (is_last) peer_signal_and_wait(pt, world, rank, channel, slot, epoch);
}

// ---- gradient averaging over peer memory ------------------------------------------------------------------------------------------
struct PeerGradTable {
  float* grad[kPeerMaxRanks];
};
__device__ __forceinline__ float4 ld_relaxed_sys_f4(const float4* p) {
  float4 v;
  asm volatile("ld.relaxed.sys.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_relaxed_sys_f4(float4* p, float4 v) {
  asm volatile("st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
// Every thread block: wait until flag[channel][0][r] in OUR workspace has reached `value` for every source rank r.
__device__ __forceinline__ void peer_wait_all(const PeerTable& pt, int world, int rank, int channel, uint32_t value) {
  if ((int)threadIdx.x < world) {
    const uint32_t* mine = reinterpret_cast<const uint32_t*>(pt.base[rank] + peer_flag_off(channel, 0, (int)threadIdx.x));
    const long long t0 = global_ns();
    while ((int32_t)(ld_acquire_sys(mine) - value) < 0) {
      if (global_ns() - t0 > kPeerWaitNs) {
        atomicExch(reinterpret_cast<unsigned int*>(pt.base[rank] + peer_err_off()), 1u);
        break;
      }
      __nanosleep(100);    // the small-block variant shares its SM with a GEMM / attention CTA: do not burn their issue slots
    }
  }
  __syncthreads();
}
// n4 float4 elements per buffer; flag values of exchange e: 2e-1 = "my gradients are final", 2e = "my slice is stored everywhere".
// Monotonic counters on one slot are safe: a rank can only signal 2e+1 after it has seen every 2e, and 2e+1 >= 2e for a late waiter.
// THREADS = 512, U = 2: one block per SM, for the exchange at the end of the backward (nothing else runs).  THREADS = 128, <= 64
// registers: blocks small enough to be co-resident with the persistent GEMM / attention CTAs (which leave ~11 K registers and 1 700
// thread slots per SM), for the exchanges issued from the side stream in the middle of the backward -- a kernel that needs whole SMs
// makes every statically scheduled persistent GEMM launched beside it wait for its last CTAs.
template <int W, int THREADS, int U>
__global__ void __launch_bounds__(THREADS, THREADS == 128 ? 8 : 1)
peer_grad_allreduce_kernel(PeerTable pt, PeerGradTable gt, long long off4, long long n4, int rank, int channel, uint32_t epoch) {
  const uint32_t ready = 2u * epoch - 1u, done = 2u * epoch;
  if (blockIdx.x == 0 && (int)threadIdx.x < W)     // stream order: every kernel that wrote this rank's gradients has completed
    st_release_sys(reinterpret_cast<uint32_t*>(pt.base[threadIdx.x] + peer_flag_off(channel, 0, rank)), ready);
  peer_wait_all(pt, W, rank, channel, ready);
  const long long per = (n4 + W - 1) / W;
  const long long lo = per * rank, hi = min(n4, lo + per);
  const float inv = 1.0f / (float)W;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = lo + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < hi; i += U * stride) {
    float4 v[U][W];
#pragma unroll
    for (int u = 0; u < U; ++u)
#pragma unroll
      for (int r = 0; r < W; ++r)
        if (i + u * stride < hi) v[u][r] = ld_relaxed_sys_f4(reinterpret_cast<const float4*>(gt.grad[r]) + off4 + i + u * stride);
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (i + u * stride >= hi) continue;
      float4 s = v[u][0];
#pragma unroll
      for (int r = 1; r < W; ++r) { s.x += v[u][r].x; s.y += v[u][r].y; s.z += v[u][r].z; s.w += v[u][r].w; }
      s.x *= inv; s.y *= inv; s.z *= inv; s.w *= inv;
#pragma unroll
      for (int r = 0; r < W; ++r) st_relaxed_sys_f4(reinterpret_cast<float4*>(gt.grad[r]) + off4 + i + u * stride, s);
    }
  }
  // last block of the grid: all slices of this rank are stored (fence + ticket) -> tell every peer, then wait for theirs
  __shared__ int is_last;
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned int* ticket = reinterpret_cast<unsigned int*>(pt.base[rank] + peer_ticket_off(channel));
    const unsigned int t = atomicAdd(ticket, 1u);
    is_last = (t == gridDim.x - 1);
    if (is_last) *ticket = 0u;
  }
  __syncthreads();
  if (!is_last) return;
  __threadfence_system();
  if ((int)threadIdx.x < W) st_release_sys(reinterpret_cast<uint32_t*>(pt.base[threadIdx.x] + peer_flag_off(channel, 0, rank)), done);
  peer_wait_all(pt, W, rank, channel, done);
}

}  // namespace dig

using namespace dig;

extern "C" int dig_peer_workspace_bytes(int64_t key_table_bytes, int64_t* total_out, int64_t* keys_offset_out) {
  DIG_REQUIRE(key_table_bytes >= 0 && total_out, "dig_peer_workspace_bytes: bad arguments");
  *total_out = (int64_t)(peer_keys_off() + 2 * (size_t)key_table_bytes);
  if (keys_offset_out) *keys_offset_out = (int64_t)peer_keys_off();
  return 0;
}

extern "C" int dig_peer_alloc(int64_t bytes, void** ptr_out, void* handle_out) {
  DIG_REQUIRE(bytes > 0 && ptr_out && handle_out, "dig_peer_alloc: bad arguments");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  void* p = nullptr;
  DIG_CHECK_CUDA(cudaMalloc(&p, (size_t)bytes));
  DIG_CHECK_CUDA(cudaMemset(p, 0, (size_t)bytes));
  DIG_CHECK_CUDA(cudaDeviceSynchronize());
  cudaIpcMemHandle_t h;
  cudaError_t e = cudaIpcGetMemHandle(&h, p);
  if (e != cudaSuccess) {
    cudaFree(p);
    set_last_error("dig_peer_alloc: cudaIpcGetMemHandle -> %s", cudaGetErrorString(e));
    return -2;
  }
  memcpy(handle_out, &h, sizeof(h));
  *ptr_out = p;
  return 0;
}

extern "C" int dig_peer_open(const void* handle, void** ptr_out) {
  DIG_REQUIRE(handle && ptr_out, "dig_peer_open: bad arguments");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, sizeof(h));
  void* p = nullptr;
  DIG_CHECK_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
  *ptr_out = p;
  return 0;
}

extern "C" int dig_peer_close(void* ptr) {
  DIG_REQUIRE(ptr != nullptr, "dig_peer_close: null pointer");
  DIG_CHECK_CUDA(cudaIpcCloseMemHandle(ptr));
  return 0;
}

extern "C" int dig_peer_free(void* ptr) {
  DIG_REQUIRE(ptr != nullptr, "dig_peer_free: null pointer");
  DIG_CHECK_CUDA(cudaFree(ptr));
  return 0;
}

namespace dig {
int fill_peer_table(PeerTable* pt, const int64_t* bases, int world) {
  DIG_REQUIRE(bases && world >= 1 && world <= kPeerMaxRanks, "peer: world size must be 1..%d", kPeerMaxRanks);
  for (int i = 0; i < kPeerMaxRanks; ++i) pt->base[i] = i < world ? reinterpret_cast<unsigned char*>((uintptr_t)bases[i]) : nullptr;
  for (int i = 0; i < world; ++i) DIG_REQUIRE(pt->base[i] != nullptr, "peer: rank %d has no mapped workspace", i);
  return 0;
}
}  // namespace dig

extern "C" int dig_peer_allreduce(const int64_t* bases, int32_t world, int32_t rank, int32_t channel, int64_t epoch, float* buf, int32_t n,
                                  void* stream) {
  DIG_REQUIRE(buf && n > 0 && n <= kPeerMaxFloats, "dig_peer_allreduce: n must be 1..%d (got %d)", kPeerMaxFloats, n);
  DIG_REQUIRE(rank >= 0 && rank < world && channel >= 0 && channel < kPeerChannels && epoch >= 1, "dig_peer_allreduce: bad rank/channel/epoch");
  PeerReduce pr;
  if (int rc = fill_peer_table(&pr.pt, bases, world)) return rc;
  pr.world = world; pr.rank = rank; pr.channel = channel; pr.epoch = (uint32_t)epoch; pr.buf = buf; pr.n = n;
  pr.loc0 = pr.loc1 = nullptr; pr.nloc = 0;
  peer_allreduce_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(pr);
  DIG_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int dig_peer_l2norm_allgather(const int64_t* bases, int32_t world, int32_t rank, int32_t channel, int64_t epoch, const float* x,
                                         float* local_copy, int64_t Q, int32_t C, int64_t key_table_bytes, void* stream) {
  DIG_REQUIRE(x && Q > 0 && C > 0, "dig_peer_l2norm_allgather: bad arguments");
  DIG_REQUIRE(rank >= 0 && rank < world && channel >= 0 && channel < kPeerChannels && epoch >= 1, "dig_peer_l2norm_allgather: bad rank/channel/epoch");
  DIG_REQUIRE((int64_t)2 * world * Q * C * (int64_t)sizeof(float) <= key_table_bytes, "dig_peer_l2norm_allgather: key table too small");
  PeerTable pt;
  if (int rc = fill_peer_table(&pt, bases, world)) return rc;
  const long long rows = 2 * Q;
  peer_l2norm_allgather_kernel<<<(int)((rows + 7) / 8), 256, 0, (cudaStream_t)stream>>>(pt, x, local_copy, Q, C, world, rank, channel,
                                                                                      (uint32_t)epoch, (size_t)key_table_bytes);
  DIG_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int dig_peer_error(const int64_t* bases, int32_t world, int32_t rank, int32_t* err_out) {
  DIG_REQUIRE(bases && err_out && rank >= 0 && rank < world && world <= kPeerMaxRanks, "dig_peer_error: bad arguments");
  uint32_t v = 0;
  DIG_CHECK_CUDA(cudaMemcpy(&v, reinterpret_cast<const unsigned char*>((uintptr_t)bases[rank]) + peer_err_off(), sizeof(v), cudaMemcpyDeviceToHost));
  *err_out = (int32_t)v;
  return 0;
}

extern "C" int dig_peer_grad_allreduce(const int64_t* bases, const int64_t* grad_bases, int32_t world, int32_t rank, int32_t channel,
                                       int64_t epoch, int64_t offset, int64_t n, int32_t blocks, int32_t small_blocks, void* stream) {
  DIG_REQUIRE(bases && grad_bases && n > 0 && n % 4 == 0 && offset >= 0 && offset % 4 == 0,
              "dig_peer_grad_allreduce: offset and n must be multiples of 4 floats (got %lld, %lld)", (long long)offset, (long long)n);
  DIG_REQUIRE(rank >= 0 && rank < world && channel >= 0 && channel < kPeerChannels && epoch >= 1 && epoch < (1ll << 30),
              "dig_peer_grad_allreduce: bad rank/channel/epoch");
  PeerTable pt;
  if (int rc = fill_peer_table(&pt, bases, world)) return rc;
  PeerGradTable gt;
  for (int i = 0; i < kPeerMaxRanks; ++i) gt.grad[i] = i < world ? reinterpret_cast<float*>((uintptr_t)grad_bases[i]) : nullptr;
  for (int i = 0; i < world; ++i) DIG_REQUIRE(gt.grad[i] != nullptr && ((uintptr_t)gt.grad[i] & 15) == 0, "dig_peer_grad_allreduce: rank %d has no (16-byte aligned) mapped gradient buffer", i);
  if (blocks <= 0) blocks = num_sms();
  const long long n4 = n / 4, off4 = offset / 4;
  cudaStream_t s = (cudaStream_t)stream;
  const uint32_t e = (uint32_t)epoch;
#define DIG_PG(W, T, U) peer_grad_allreduce_kernel<W, T, U><<<blocks, T, 0, s>>>(pt, gt, off4, n4, rank, channel, e)
  switch (world * 2 + (small_blocks ? 1 : 0)) {
    case 4: DIG_PG(2, 512, 2); break;
    case 5: DIG_PG(2, 128, 4); break;
    case 8: DIG_PG(4, 512, 2); break;
    case 9: DIG_PG(4, 128, 2); break;
    case 16: DIG_PG(8, 512, 2); break;
    case 17: DIG_PG(8, 128, 1); break;
    default:
      set_last_error("dig_peer_grad_allreduce: built for 2, 4 or 8 ranks (got %d)", world);
      return -1;
  }
#undef DIG_PG
  DIG_CHECK_CUDA(cudaGetLastError());
  return 0;
}
