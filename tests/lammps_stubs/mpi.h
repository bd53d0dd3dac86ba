// single-rank MPI: every collective is a copy
#pragma once
#include <chrono>
#include <cstring>
typedef int MPI_Comm;
typedef int MPI_Datatype;
typedef int MPI_Op;
typedef int MPI_Info;
#define MPI_INT 4
#define MPI_DOUBLE 8
#define MPI_BYTE 1
#define MPI_SUM 1
#define MPI_MAX 2
#define MPI_MIN 3
#define MPI_COMM_WORLD 0
#define MPI_COMM_TYPE_SHARED 1
#define MPI_INFO_NULL 0
inline int MPI_Bcast(void *, int, MPI_Datatype, int, MPI_Comm) { return 0; }
inline int MPI_Allreduce(const void *s, void *r, int n, MPI_Datatype t, MPI_Op, MPI_Comm) { memcpy(r, s, (size_t)n * t); return 0; }
inline int MPI_Reduce(const void *s, void *r, int n, MPI_Datatype t, MPI_Op, int, MPI_Comm) { memcpy(r, s, (size_t)n * t); return 0; }
inline int MPI_Allgather(const void *s, int n, MPI_Datatype t, void *r, int, MPI_Datatype, MPI_Comm) { memcpy(r, s, (size_t)n * t); return 0; }
inline int MPI_Allgatherv(const void *s, int n, MPI_Datatype t, void *r, const int *, const int *d, MPI_Datatype, MPI_Comm) {
  memcpy((char *)r + (size_t)d[0] * t, s, (size_t)n * t);
  return 0;
}
inline int MPI_Gatherv(const void *s, int n, MPI_Datatype t, void *r, const int *, const int *d, MPI_Datatype, int, MPI_Comm) {
  memcpy((char *)r + (size_t)d[0] * t, s, (size_t)n * t);
  return 0;
}
inline int MPI_Barrier(MPI_Comm) { return 0; }
inline int MPI_Comm_split_type(MPI_Comm c, int, int, MPI_Info, MPI_Comm *out) { *out = c; return 0; }
inline int MPI_Comm_rank(MPI_Comm, int *r) { *r = 0; return 0; }
inline int MPI_Comm_free(MPI_Comm *) { return 0; }
inline double MPI_Wtime() {
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
