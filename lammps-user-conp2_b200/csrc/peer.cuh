// Device side of the peer-to-peer exchanges (comm.cu): the control block every rank keeps at the
// start of its IPC-mapped arena, and the helpers that let a compute kernel do its own exchange --
// store results straight into every peer's arena and raise the flags from its last block
// (peer_block_signal), or poll the flags in its prologue instead of in a separate wait kernel
// (peer_block_wait).  Replaces MPI_Allgatherv / MPI_Allreduce of fix_conp.cpp:641-648, 1140 and
// km_ewald.cpp:842 on the step's latency-critical small vectors.
//
// Protocol: channel `chan` carries one exchange per step.  Rank r announces its contribution by
// writing the epoch number e = epoch[chan] + 1 into flags[chan][r] of every peer's control block
// (st.release.sys after a system-scope fence); a consumer spins until all flags of the channel have
// reached e and then advances its local epoch[chan].  All ranks run the same sequence of exchanges,
// so the epochs stay in lock-step.
#pragma once
#include <cstdint>

namespace conp {

constexpr int P2P_CHANNELS = 8;
constexpr int P2P_MAX_RANKS = 16;
constexpr size_t P2P_CTRL_BYTES = 4096;

struct ArenaCtl {
  unsigned long long flags[P2P_CHANNELS][P2P_MAX_RANKS];  // written by the peers
  unsigned long long epoch[P2P_CHANNELS];                 // local
  int error;                                              // a wait timed out
  unsigned int prod_ticket[P2P_CHANNELS];  // blocks of a fused producer kernel that have stored their share
  unsigned int cons_ticket[P2P_CHANNELS];  // blocks of a fused consumer kernel that have seen the flags
};

// by-value kernel argument; arena == nullptr: single GPU / NCCL path, every helper is a no-op
struct PeerSync {
  char *const *arena = nullptr;  // device array: mapped base of every rank's arena
  int rank = 0, nranks = 1, chan = 0;
  // 1: the producer kernel raises the flags itself (one system fence per block); 0: it only stores, and
  // a one-block p2p_signal kernel raises them after the kernel boundary (cheaper for grids of many
  // small blocks: measured ~20 us of fences per kernel at ~1000 blocks)
  int signal_in_kernel = 1;
  // consumer side: 1 = this rank's flags of the channel have not been raised yet (the producer kernel before
  // this one only stored); block (0,0) of the consumer raises them in peer_block_wait before it starts to
  // poll -- the kernel boundary has completed the producer's stores, and no stand-alone signal kernel is needed
  int raise_first = 0;
  // payload that only exists once the whole producer grid is done, delivered with the flags (raise_first):
  // *value -> byte offset value_off of every rank's arena; counts[r] -> slot `rank` at count_off of rank r's
  size_t value_off = 0, count_off = 0;
  const double *value = nullptr;
  const int *counts = nullptr;
};

__device__ __forceinline__ unsigned long long peer_ld_acquire(const unsigned long long *p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void peer_st_release(unsigned long long *p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// payload address `off` bytes into rank r's arena
template <class T>
__device__ __forceinline__ T *peer_ptr(const PeerSync &ps, int r, size_t off) {
  return reinterpret_cast<T *>(ps.arena[r] + P2P_CTRL_BYTES + off);
}

// Producer side, called by ALL threads of every block after the block's stores into the peers' arenas
// (contains __syncthreads).  The last block to arrive raises this rank's flag on every peer.
// `before_flags(last)` runs in the last block (all threads) before the flags go up, for a final
// payload that needs the whole grid's result.
template <class F>
__device__ __forceinline__ void peer_block_signal(const PeerSync &ps, F &&before_flags) {
  if (!ps.arena || !ps.signal_in_kernel) return;
  __shared__ int s_last;
  ArenaCtl *mine = reinterpret_cast<ArenaCtl *>(ps.arena[ps.rank]);
  // The barrier orders the block's stores before thread 0's fence, and fences are cumulative: one
  // system-scope fence per block (not per thread) publishes them before the ticket is taken.
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence_system();
    const unsigned int t = atomicAdd(&mine->prod_ticket[ps.chan], 1u);
    s_last = (t == gridDim.x * gridDim.y - 1);
    if (s_last) {
      mine->prod_ticket[ps.chan] = 0u;
      __threadfence();  // acquire side: see what the other blocks stored / accumulated
    }
  }
  __syncthreads();
  if (!s_last) return;
  before_flags();
  __syncthreads();
  if ((int)threadIdx.x < ps.nranks && (int)threadIdx.x != ps.rank) {
    const unsigned long long e = mine->epoch[ps.chan] + 1ull;
    ArenaCtl *theirs = reinterpret_cast<ArenaCtl *>(ps.arena[threadIdx.x]);
    __threadfence_system();
    peer_st_release(&theirs->flags[ps.chan][ps.rank], e);
  }
}
__device__ __forceinline__ void peer_block_signal(const PeerSync &ps) {
  peer_block_signal(ps, [] {});
}

// Consumer side, called by the threads 0 .. nranks-1 (at least) of every block before the block reads
// exchanged data; the caller then synchronises the threads that will read.  The last block through
// advances the local epoch.  Requires blockDim.x >= nranks.
__device__ __forceinline__ void peer_block_wait(const PeerSync &ps) {
  if (!ps.arena) return;
  ArenaCtl *mine = reinterpret_cast<ArenaCtl *>(ps.arena[ps.rank]);
  const unsigned long long e = mine->epoch[ps.chan] + 1ull;
  const int r = threadIdx.x;
  if (ps.raise_first && blockIdx.x == 0 && blockIdx.y == 0 && r < ps.nranks) {
    // (blocks are dispatched in index order, so block (0,0) of every rank runs even if the later blocks of
    // the grid fill the GPU with pollers)
    if (ps.value) *reinterpret_cast<double *>(ps.arena[r] + ps.value_off) = __ldcg(ps.value);
    if (ps.counts) reinterpret_cast<int *>(ps.arena[r] + ps.count_off)[ps.rank] = __ldcg(ps.counts + r);
    if (r != ps.rank) {
      __threadfence_system();
      peer_st_release(&reinterpret_cast<ArenaCtl *>(ps.arena[r])->flags[ps.chan][ps.rank], e);
    }
  }
  if (r < ps.nranks && r != ps.rank) {
    const long long t0 = clock64();
    while (peer_ld_acquire(&mine->flags[ps.chan][r]) < e) {
      if (clock64() - t0 > 20000000000ll) {  // ~10 s: a peer is gone; fail instead of hanging the GPU
        mine->error = 1;
        break;
      }
    }
  }
  __syncwarp();
  if (r == 0) {
    const unsigned int t = atomicAdd(&mine->cons_ticket[ps.chan], 1u);
    if (t == gridDim.x * gridDim.y - 1) {  // every block has read the epoch: move on
      mine->cons_ticket[ps.chan] = 0u;
      mine->epoch[ps.chan] = e;
    }
  }
}

}  // namespace conp
