// NCCL plumbing for the row-sharded multi-GPU path.  Replaces the MPI
// collectives of the reference on LAMMPS' `world` communicator:
//   b_comm  (MPI_Allgatherv + permutation)   fix_conp.cpp:641-648  -> ncclAllGather of equal row blocks
//   sfac_reduce / scalar Allreduce            km_ewald.cpp:782-786, 842   -> ncclAllReduce
// NCCL is resolved at run time with dlopen so that the library binds to the
// libnccl already loaded in the process (torch's bundled copy in the Python
// harness, the system copy under LAMMPS) instead of pulling in a second one.
#include "common.cuh"

#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <cstddef>
#include <cstdlib>
#include <cstring>
#include <vector>

namespace conp {

namespace {

struct NcclApi {
  void *handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t,
                            cudaStream_t) = nullptr;
  ncclResult_t (*Broadcast)(const void *, void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char *(*GetErrorString)(ncclResult_t) = nullptr;
};

NcclApi &api() {
  static NcclApi a;
  if (a.handle) return a;
  const char *names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char *n : names) {
    a.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (a.handle) break;
  }
  if (!a.handle) CONP_THROW(CONP_ERR_COMM, "cannot dlopen libnccl.so.2: %s", dlerror());
#define LOAD(field, sym)                                                       \
  *(void **)(&a.field) = dlsym(a.handle, sym);                                 \
  if (!a.field) CONP_THROW(CONP_ERR_COMM, "libnccl lacks symbol %s", sym)
  LOAD(GetUniqueId, "ncclGetUniqueId");
  LOAD(CommInitRank, "ncclCommInitRank");
  LOAD(CommDestroy, "ncclCommDestroy");
  LOAD(AllGather, "ncclAllGather");
  LOAD(AllReduce, "ncclAllReduce");
  LOAD(Broadcast, "ncclBroadcast");
  LOAD(GroupStart, "ncclGroupStart");
  LOAD(GroupEnd, "ncclGroupEnd");
  LOAD(GetErrorString, "ncclGetErrorString");
#undef LOAD
  return a;
}

#define NCCL_CHECK(expr)                                                                          \
  do {                                                                                            \
    ncclResult_t r_ = (expr);                                                                     \
    if (r_ != ncclSuccess)                                                                        \
      CONP_THROW(CONP_ERR_COMM, "NCCL error %s at %s:%d (%s)", api().GetErrorString(r_), __FILE__, \
                 __LINE__, #expr);                                                                \
  } while (0)

}  // namespace

struct Comm {
  ncclComm_t comm = nullptr;
  int rank = 0, nranks = 1;
};

static_assert(sizeof(ncclUniqueId) <= CONP_UNIQUE_ID_BYTES, "unique id does not fit the ABI buffer");

int comm_get_unique_id(void *out) {
  ncclUniqueId id;
  NCCL_CHECK(api().GetUniqueId(&id));
  std::memset(out, 0, CONP_UNIQUE_ID_BYTES);
  std::memcpy(out, &id, sizeof(id));
  return 0;
}

Comm *comm_create(int rank, int nranks, const void *unique_id) {
  Comm *c = new Comm;
  c->rank = rank;
  c->nranks = nranks;
  if (nranks > 1) {
    if (!unique_id) {
      delete c;
      CONP_THROW(CONP_ERR_ARG, "conp_create: nranks > 1 needs a unique id");
    }
    ncclUniqueId id;
    std::memcpy(&id, unique_id, sizeof(id));
    ncclResult_t r = api().CommInitRank(&c->comm, nranks, id, rank);
    if (r != ncclSuccess) {
      delete c;
      CONP_THROW(CONP_ERR_COMM, "ncclCommInitRank failed: %s", api().GetErrorString(r));
    }
  }
  return c;
}

void comm_destroy(Comm *c) {
  if (!c) return;
  if (c->comm) api().CommDestroy(c->comm);
  delete c;
}

void comm_allgather(Comm *c, const void *send, void *recv, size_t bytes_per_rank, cudaStream_t s) {
  if (!c || c->nranks == 1) return;
  NCCL_CHECK(api().AllGather(send, recv, bytes_per_rank, ncclChar, c->comm, s));
}

void comm_allgatherv(Comm *c, const void *send, void *recv, const size_t *bytes, const size_t *offsets, int rank,
                     int nranks, cudaStream_t s) {
  if (!c || c->nranks == 1) return;
  (void)rank;
  NCCL_CHECK(api().GroupStart());
  for (int r = 0; r < nranks; ++r) {
    if (bytes[r] == 0) continue;
    void *dst = (char *)recv + offsets[r];
    const void *src = (r == c->rank) ? send : dst;
    NCCL_CHECK(api().Broadcast(src, dst, bytes[r], ncclChar, r, c->comm, s));
  }
  NCCL_CHECK(api().GroupEnd());
}

void comm_allreduce_sum_f64(Comm *c, double *buf, size_t n, cudaStream_t s) {
  if (!c || c->nranks == 1) return;
  NCCL_CHECK(api().AllReduce(buf, buf, n, ncclDouble, ncclSum, c->comm, s));
}

// ===========================================================================
// direct peer-to-peer exchanges
// ===========================================================================
struct PeerArena {
  Comm *comm = nullptr;
  int rank = 0, nranks = 1;
  size_t bytes = 0;
  char *local = nullptr;
  std::vector<char *> peer;        // mapped base of every rank's arena (peer[rank] == local)
  char **d_peer = nullptr;         // device copy
  // control block at the start of every arena: flags[chan][rank] (written by peers), then local-only
  // epoch[chan] and an error word
  static constexpr size_t CTRL_BYTES = P2P_CTRL_BYTES;
};

namespace {

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p) { return peer_ld_acquire(p); }
__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v) { peer_st_release(p, v); }

static_assert(sizeof(ArenaCtl) + P2P_CHANNELS * 16 * sizeof(unsigned int) <= P2P_CTRL_BYTES, "control block too large");

// grid = (blocks_per_peer, nranks-1); peer index p = (rank + 1 + blockIdx.y) % nranks
__global__ void __launch_bounds__(256)
p2p_push_kernel(char *const *__restrict__ arena, int rank, int nranks, size_t off, size_t bytes, int chan) {
  const int p = (rank + 1 + blockIdx.y) % nranks;
  const uint4 *src = reinterpret_cast<const uint4 *>(arena[rank] + off);
  uint4 *dst = reinterpret_cast<uint4 *>(arena[p] + off);
  const size_t n16 = bytes / 16;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x)
    dst[i] = src[i];
  // last block of this peer's column raises the flag once all of the column's stores are out
  __shared__ bool last;
  __threadfence_system();
  __syncthreads();
  ArenaCtl *mine = reinterpret_cast<ArenaCtl *>(arena[rank]);
  if (threadIdx.x == 0) {
    // per-peer arrival counter lives in the (local-only) tail of my control block
    unsigned int *cnt = reinterpret_cast<unsigned int *>(arena[rank] + sizeof(ArenaCtl)) + chan * 16 + p;
    const unsigned int t = atomicAdd(cnt, 1u);
    last = (t == gridDim.x - 1);
    if (last) *cnt = 0u;
  }
  __syncthreads();
  if (last && threadIdx.x == 0) {
    const unsigned long long e = mine->epoch[chan] + 1ull;  // bumped by the wait kernel that follows
    ArenaCtl *theirs = reinterpret_cast<ArenaCtl *>(arena[p]);
    __threadfence_system();
    st_release_sys(&theirs->flags[chan][rank], e);
  }
}

__global__ void __launch_bounds__(32)
p2p_wait_kernel(char *const *__restrict__ arena, int rank, int nranks, int chan) {
  ArenaCtl *mine = reinterpret_cast<ArenaCtl *>(arena[rank]);
  const unsigned long long e = mine->epoch[chan] + 1ull;
  const int r = threadIdx.x;
  if (r < nranks && r != rank) {
    const long long t0 = clock64();
    while (ld_acquire_sys(&mine->flags[chan][r]) < e) {
      if (clock64() - t0 > 20000000000ll) {  // ~10 s: a peer is gone; fail instead of hanging the GPU
        mine->error = 1;
        break;
      }
    }
  }
  __syncwarp();
  if (r == 0) mine->epoch[chan] = e;
}

// reduce-scatter stage 2 + all-gather: I own slice `rank` of the vector.  Sum the nranks partial
// slices (mine from the vector itself, the others from the staging area, in rank order so the result
// is deterministic), store it into my vector and into every peer's vector, then signal chan.
__global__ void __launch_bounds__(256)
p2p_reduce_bcast_kernel(char *const *__restrict__ arena, int rank, int nranks, size_t off, size_t n, size_t slice,
                        size_t stage_off, int chan) {
  const size_t lo = (size_t)rank * slice;
  const size_t cnt = lo < n ? (n - lo < slice ? n - lo : slice) : 0;
  double *vec = reinterpret_cast<double *>(arena[rank] + off);
  const double *stage = reinterpret_cast<const double *>(arena[rank] + stage_off);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < cnt; i += (size_t)gridDim.x * blockDim.x) {
    double v = 0.0;
    for (int r = 0; r < nranks; ++r) v += (r == rank) ? vec[lo + i] : stage[(size_t)r * slice + i];
    for (int p = 0; p < nranks; ++p) reinterpret_cast<double *>(arena[p] + off)[lo + i] = v;
  }
  __shared__ bool last;
  __threadfence_system();
  __syncthreads();
  ArenaCtl *mine = reinterpret_cast<ArenaCtl *>(arena[rank]);
  if (threadIdx.x == 0) {
    unsigned int *c2 = reinterpret_cast<unsigned int *>(arena[rank] + sizeof(ArenaCtl)) + chan * 16 + 15;
    const unsigned int t = atomicAdd(c2, 1u);
    last = (t == gridDim.x - 1);
    if (last) *c2 = 0u;
  }
  __syncthreads();
  if (last && threadIdx.x < nranks && (int)threadIdx.x != rank) {
    const unsigned long long e = mine->epoch[chan] + 1ull;
    ArenaCtl *theirs = reinterpret_cast<ArenaCtl *>(arena[threadIdx.x]);
    __threadfence_system();
    st_release_sys(&theirs->flags[chan][rank], e);
  }
}

// Raise this rank's flag of `chan` on every peer: launched right after a producer kernel that stored
// into the peers' arenas without signalling (PeerSync::signal_in_kernel == 0).  The kernel boundary has
// completed those stores; optionally one double (*value) is first copied to byte offset value_off of
// every rank's arena (the rank's sum(q z), which only exists once the whole producer grid is done).
__global__ void __launch_bounds__(32)
p2p_signal_kernel(PeerSync ps, size_t value_off, const double *__restrict__ value, size_t count_off,
                  const int *__restrict__ counts) {
  const int r = threadIdx.x;
  if (r >= ps.nranks) return;
  ArenaCtl *mine = reinterpret_cast<ArenaCtl *>(ps.arena[ps.rank]);
  if (value) *reinterpret_cast<double *>(ps.arena[r] + value_off) = __ldcg(value);
  if (counts) reinterpret_cast<int *>(ps.arena[r] + count_off)[ps.rank] = __ldcg(counts + r);
  if (r == ps.rank) return;
  const unsigned long long e = mine->epoch[ps.chan] + 1ull;
  __threadfence_system();
  peer_st_release(&reinterpret_cast<ArenaCtl *>(ps.arena[r])->flags[ps.chan][ps.rank], e);
}

// Pull variant of the all-reduce for a vector whose producer kernel has already raised the flags of
// `ps_ready` on every peer (peer_block_signal in its tail): I own slice `rank`; wait for everybody's
// partial, read the slice from every rank's copy of the vector (mine locally, the others over NVLink,
// in rank order so that the sum is deterministic), and store the total into every rank's vector.  One
// kernel and one flag round instead of scatter + wait + reduce.
__global__ void __launch_bounds__(256)
p2p_pull_reduce_bcast_kernel(PeerSync ps_ready, PeerSync ps_done, size_t off, size_t n, size_t slice) {
  peer_block_wait(ps_ready);
  __syncthreads();
  const int rank = ps_ready.rank, nranks = ps_ready.nranks;
  const size_t lo = (size_t)rank * slice;
  const size_t cnt = lo < n ? (n - lo < slice ? n - lo : slice) : 0;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < cnt; i += (size_t)gridDim.x * blockDim.x) {
    double part[P2P_MAX_RANKS];
#pragma unroll
    for (int r = 0; r < P2P_MAX_RANKS; ++r)
      if (r < nranks) part[r] = __ldcg(reinterpret_cast<const double *>(ps_ready.arena[r] + off) + lo + i);
    double v = 0.0;
#pragma unroll
    for (int r = 0; r < P2P_MAX_RANKS; ++r)
      if (r < nranks) v += part[r];
#pragma unroll
    for (int r = 0; r < P2P_MAX_RANKS; ++r)
      if (r < nranks) reinterpret_cast<double *>(ps_ready.arena[r] + off)[lo + i] = v;
  }
  peer_block_signal(ps_done);
}

// reduce-scatter stage 1: send slice o of my partial vector to owner o's staging row `rank`
__global__ void __launch_bounds__(256)
p2p_scatter_kernel(char *const *__restrict__ arena, int rank, int nranks, size_t off, size_t n, size_t slice,
                   size_t stage_off, int chan) {
  const int o = (rank + 1 + blockIdx.y) % nranks;
  const size_t lo = (size_t)o * slice;
  const size_t cnt = lo < n ? (n - lo < slice ? n - lo : slice) : 0;
  const double *vec = reinterpret_cast<const double *>(arena[rank] + off);
  double *dst = reinterpret_cast<double *>(arena[o] + stage_off) + (size_t)rank * slice;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < cnt; i += (size_t)gridDim.x * blockDim.x)
    dst[i] = vec[lo + i];
  __shared__ bool last;
  __threadfence_system();
  __syncthreads();
  ArenaCtl *mine = reinterpret_cast<ArenaCtl *>(arena[rank]);
  if (threadIdx.x == 0) {
    unsigned int *c2 = reinterpret_cast<unsigned int *>(arena[rank] + sizeof(ArenaCtl)) + chan * 16 + o;
    const unsigned int t = atomicAdd(c2, 1u);
    last = (t == gridDim.x - 1);
    if (last) *c2 = 0u;
  }
  __syncthreads();
  if (last && threadIdx.x == 0) {
    const unsigned long long e = mine->epoch[chan] + 1ull;
    ArenaCtl *theirs = reinterpret_cast<ArenaCtl *>(arena[o]);
    __threadfence_system();
    st_release_sys(&theirs->flags[chan][rank], e);
  }
}

}  // namespace

PeerArena *p2p_create(Comm *c, size_t bytes, cudaStream_t s) {
  if (!c || c->nranks == 1 || c->nranks > 16) return nullptr;
  if (getenv("CONP_P2P") && atoi(getenv("CONP_P2P")) == 0) return nullptr;
  PeerArena *a = new PeerArena;
  a->comm = c; a->rank = c->rank; a->nranks = c->nranks;
  a->bytes = (bytes + PeerArena::CTRL_BYTES + 255) / 256 * 256;
  int ok = 1;
  cudaIpcMemHandle_t mine;
  if (cudaMalloc((void **)&a->local, a->bytes) != cudaSuccess) ok = 0;
  if (ok && cudaMemsetAsync(a->local, 0, a->bytes, s) != cudaSuccess) ok = 0;
  if (ok && cudaIpcGetMemHandle(&mine, a->local) != cudaSuccess) ok = 0;
  // exchange the handles (and everybody's ok flag) through NCCL
  const size_t rec = sizeof(cudaIpcMemHandle_t) + 8;
  std::vector<char> h((size_t)c->nranks * rec, 0);
  char *d = nullptr;
  cudaMalloc((void **)&d, h.size());
  std::memcpy(h.data() + (size_t)c->rank * rec, &mine, sizeof(mine));
  h[(size_t)c->rank * rec + sizeof(mine)] = (char)ok;
  cudaMemcpyAsync(d + (size_t)c->rank * rec, h.data() + (size_t)c->rank * rec, rec, cudaMemcpyHostToDevice, s);
  comm_allgather(c, d + (size_t)c->rank * rec, d, rec, s);
  cudaMemcpyAsync(h.data(), d, h.size(), cudaMemcpyDeviceToHost, s);
  cudaStreamSynchronize(s);
  cudaFree(d);
  for (int r = 0; r < c->nranks; ++r) ok = ok && h[(size_t)r * rec + sizeof(mine)];
  a->peer.assign(c->nranks, nullptr);
  if (ok) {
    for (int r = 0; r < c->nranks && ok; ++r) {
      if (r == c->rank) { a->peer[r] = a->local; continue; }
      cudaIpcMemHandle_t hd;
      std::memcpy(&hd, h.data() + (size_t)r * rec, sizeof(hd));
      void *ptr = nullptr;
      if (cudaIpcOpenMemHandle(&ptr, hd, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { ok = 0; cudaGetLastError(); }
      a->peer[r] = (char *)ptr;
    }
  }
  // agree on the outcome: everybody falls back to NCCL unless every mapping worked
  {
    double *flag = nullptr;
    cudaMalloc((void **)&flag, sizeof(double));
    const double v = ok ? 0.0 : 1.0;
    cudaMemcpyAsync(flag, &v, sizeof(double), cudaMemcpyHostToDevice, s);
    comm_allreduce_sum_f64(c, flag, 1, s);
    double tot = 0;
    cudaMemcpyAsync(&tot, flag, sizeof(double), cudaMemcpyDeviceToHost, s);
    cudaStreamSynchronize(s);
    cudaFree(flag);
    if (tot != 0.0) ok = 0;
  }
  if (!ok) {
    p2p_destroy(a);
    return nullptr;
  }
  cudaMalloc((void **)&a->d_peer, sizeof(char *) * c->nranks);
  cudaMemcpyAsync(a->d_peer, a->peer.data(), sizeof(char *) * c->nranks, cudaMemcpyHostToDevice, s);
  cudaStreamSynchronize(s);
  return a;
}

void p2p_destroy(PeerArena *a) {
  if (!a) return;
  for (int r = 0; r < (int)a->peer.size(); ++r)
    if (r != a->rank && a->peer[r]) cudaIpcCloseMemHandle(a->peer[r]);
  if (a->d_peer) cudaFree(a->d_peer);
  if (a->local) cudaFree(a->local);
  delete a;
}

char *p2p_local(PeerArena *a) { return a->local + PeerArena::CTRL_BYTES; }

static inline size_t arena_off(size_t off) { return off + PeerArena::CTRL_BYTES; }

int p2p_push(PeerArena *a, size_t off, size_t bytes, int chan, cudaStream_t s) {
  if (bytes % 16) CONP_THROW(CONP_ERR_ARG, "p2p_push: size must be a multiple of 16 bytes");
  unsigned bx = (unsigned)std::min<size_t>(32, std::max<size_t>(1, bytes / (16 * 256 * 4)));
  dim3 grid(bx, a->nranks - 1);
  p2p_push_kernel<<<grid, 256, 0, s>>>(a->d_peer, a->rank, a->nranks, arena_off(off), bytes, chan);
  CUDA_CHECK(cudaGetLastError());
  return 1;
}

int p2p_wait(PeerArena *a, int chan, cudaStream_t s) {
  p2p_wait_kernel<<<1, 32, 0, s>>>(a->d_peer, a->rank, a->nranks, chan);
  CUDA_CHECK(cudaGetLastError());
  return 1;
}

int p2p_wait_sync(const PeerSync &ps, cudaStream_t s) {
  if (!ps.arena) return 0;
  p2p_wait_kernel<<<1, 32, 0, s>>>(ps.arena, ps.rank, ps.nranks, ps.chan);
  CUDA_CHECK(cudaGetLastError());
  return 1;
}

int p2p_allgather(PeerArena *a, size_t off, size_t stride_bytes, size_t block_bytes, int chan, cudaStream_t s) {
  int n = p2p_push(a, off + (size_t)a->rank * stride_bytes, block_bytes, chan, s);
  return n + p2p_wait(a, chan, s);
}

int p2p_allreduce_f64(PeerArena *a, size_t off, size_t n, size_t stage_off, int chan, cudaStream_t s) {
  const size_t slice = (n + a->nranks - 1) / a->nranks;
  unsigned bx = (unsigned)std::min<size_t>(32, std::max<size_t>(1, slice / (256 * 4)));
  dim3 grid(bx, a->nranks - 1);
  p2p_scatter_kernel<<<grid, 256, 0, s>>>(a->d_peer, a->rank, a->nranks, arena_off(off), n, slice,
                                          arena_off(stage_off), chan);
  CUDA_CHECK(cudaGetLastError());
  p2p_wait(a, chan, s);
  p2p_reduce_bcast_kernel<<<bx, 256, 0, s>>>(a->d_peer, a->rank, a->nranks, arena_off(off), n, slice,
                                             arena_off(stage_off), chan + 1);
  CUDA_CHECK(cudaGetLastError());
  p2p_wait(a, chan + 1, s);
  return 4;
}

PeerSync p2p_sync(PeerArena *a, int chan) {
  PeerSync ps;
  if (!a) return ps;
  ps.arena = a->d_peer;
  ps.rank = a->rank;
  ps.nranks = a->nranks;
  ps.chan = chan;
  return ps;
}

size_t p2p_arena_offset(size_t payload_off) { return arena_off(payload_off); }

int p2p_signal(PeerArena *a, int chan, cudaStream_t s, size_t value_off, const double *value, size_t count_off,
               const int *counts) {
  p2p_signal_kernel<<<1, 32, 0, s>>>(p2p_sync(a, chan), value ? arena_off(value_off) : 0, value,
                                     counts ? arena_off(count_off) : 0, counts);
  CUDA_CHECK(cudaGetLastError());
  return 1;
}

int p2p_allreduce_pull_f64(PeerArena *a, size_t off, size_t n, int chan_ready, int chan_done, cudaStream_t s,
                           int raise_ready) {
  const size_t slice = (n + a->nranks - 1) / a->nranks;
  const unsigned grid = (unsigned)std::min<size_t>(296, std::max<size_t>(1, (slice + 255) / 256));
  PeerSync ready = p2p_sync(a, chan_ready);
  ready.raise_first = raise_ready;
  p2p_pull_reduce_bcast_kernel<<<grid, 256, 0, s>>>(ready, p2p_sync(a, chan_done), arena_off(off), n, slice);
  CUDA_CHECK(cudaGetLastError());
  return 1 + p2p_wait(a, chan_done, s);
}

int p2p_error(PeerArena *a) {
  int e = 0;
  cudaMemcpy(&e, a->local + offsetof(ArenaCtl, error), sizeof(int), cudaMemcpyDeviceToHost);
  return e;
}

}  // namespace conp
