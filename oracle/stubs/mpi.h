/* Single-rank MPI stand-in used ONLY to compile the unmodified reference sources
 * (/root/reference/dmc/*.cpp) into oracle/_ref.  TEST INFRASTRUCTURE, not product code.
 *
 * The reference touches MPI in dmc/ptools.cpp, dmc/walkers.cpp and dmc/walkers.h:70 only.
 * With one rank: size=1, rank=0, every reduction/gather is a copy, and the alias-method
 * planner (ptools.cpp:93-161) never schedules a transfer, so point-to-point calls are
 * unreachable (they abort if reached).
 */
#ifndef QMC4_ORACLE_MPI_STUB_H
#define QMC4_ORACLE_MPI_STUB_H
#include <cstddef>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>

typedef int MPI_Comm;
typedef int MPI_Datatype;
typedef int MPI_Op;
typedef int MPI_Request;
typedef std::intptr_t MPI_Aint;
struct MPI_Status { int MPI_SOURCE, MPI_TAG, MPI_ERROR; };

#define MPI_SUCCESS 0
#define MPI_COMM_WORLD 0
#define MPI_SUM 1
#define MPI_BOTTOM ((void*)0)
#define MPI_IN_PLACE ((void*)1)

/* datatype handle = element size in bytes (enough for memcpy-collectives) */
#define MPI_CHAR 1
#define MPI_UNSIGNED_CHAR 1
#define MPI_CXX_BOOL 1
#define MPI_UNSIGNED_SHORT 2
#define MPI_INT 4
#define MPI_UNSIGNED 4
#define MPI_DOUBLE 8
#define MPI_UNSIGNED_LONG 8
#define MPI_UNSIGNED_LONG_LONG 8
#define MPI_STRUCT_STUB 0

static inline int MPI_Init(int*, char***) { return MPI_SUCCESS; }
static inline int MPI_Finalize() { return MPI_SUCCESS; }
static inline int MPI_Comm_size(MPI_Comm, int* n) { *n = 1; return MPI_SUCCESS; }
static inline int MPI_Comm_rank(MPI_Comm, int* r) { *r = 0; return MPI_SUCCESS; }
static inline int MPI_Barrier(MPI_Comm) { return MPI_SUCCESS; }
static inline int MPI_Bcast(void*, int, MPI_Datatype, int, MPI_Comm) { return MPI_SUCCESS; }

static inline int mpi_stub_copy(const void* s, void* r, int count, MPI_Datatype t)
{
  if (s != MPI_IN_PLACE && s != r) std::memcpy(r, s, (size_t)count * (size_t)t);
  return MPI_SUCCESS;
}
static inline int MPI_Reduce(const void* s, void* r, int c, MPI_Datatype t, MPI_Op, int, MPI_Comm)
{ return mpi_stub_copy(s, r, c, t); }
static inline int MPI_Allreduce(const void* s, void* r, int c, MPI_Datatype t, MPI_Op, MPI_Comm)
{ return mpi_stub_copy(s, r, c, t); }
static inline int MPI_Allgather(const void* s, int sc, MPI_Datatype st, void* r, int, MPI_Datatype, MPI_Comm)
{ return mpi_stub_copy(s, r, sc, st); }
static inline int MPI_Gather(const void* s, int sc, MPI_Datatype st, void* r, int, MPI_Datatype, int, MPI_Comm)
{ return mpi_stub_copy(s, r, sc, st); }
static inline int MPI_Gatherv(const void* s, int sc, MPI_Datatype st, void* r, const int*, const int* displs,
                              MPI_Datatype rt, int, MPI_Comm)
{ return mpi_stub_copy(s, (char*)r + (displs ? (size_t)displs[0] * (size_t)rt : 0), sc, st); }

static inline int mpi_stub_unreachable(const char* what)
{
  std::fprintf(stderr, "mpi stub: %s reached in a single-rank build\n", what);
  std::abort();
  return 1;
}
static inline int MPI_Isend(const void*, int, MPI_Datatype, int, int, MPI_Comm, MPI_Request*) { return mpi_stub_unreachable("MPI_Isend"); }
static inline int MPI_Irecv(void*, int, MPI_Datatype, int, int, MPI_Comm, MPI_Request*) { return mpi_stub_unreachable("MPI_Irecv"); }
static inline int MPI_Send(const void*, int, MPI_Datatype, int, int, MPI_Comm) { return mpi_stub_unreachable("MPI_Send"); }
static inline int MPI_Recv(void*, int, MPI_Datatype, int, int, MPI_Comm, MPI_Status*) { return mpi_stub_unreachable("MPI_Recv"); }
static inline int MPI_Wait(MPI_Request*, MPI_Status*) { return MPI_SUCCESS; }
static inline int MPI_Waitall(int, MPI_Request*, MPI_Status*) { return MPI_SUCCESS; }

static inline int MPI_Get_address(const void* p, MPI_Aint* a) { *a = (MPI_Aint)p; return MPI_SUCCESS; }
static inline int MPI_Type_create_struct(int, const int*, const MPI_Aint*, const MPI_Datatype*, MPI_Datatype* t)
{ *t = MPI_STRUCT_STUB; return MPI_SUCCESS; }
static inline int MPI_Type_commit(MPI_Datatype*) { return MPI_SUCCESS; }
static inline int MPI_Type_free(MPI_Datatype*) { return MPI_SUCCESS; }
#endif
