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

#include <cstring>

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

}  // namespace conp
