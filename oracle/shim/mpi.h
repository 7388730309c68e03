/* Single-rank MPI stand-in so the reference's main.cpp (which includes <mpi.h>
 * unconditionally, main.cpp:4) compiles in a container without an MPI toolchain.
 * TEST INFRASTRUCTURE ONLY (oracle/_ref build). Data movement only; MPI_Bcast / MPI_Allreduce / MPI_INT / MPI_BYTE /
 * MPI_MIN are here for the product's host shim (rt_render_shim.hpp), which a real MPI serves when there are several ranks. */
#ifndef ORACLE_SHIM_MPI_H_
#define ORACLE_SHIM_MPI_H_
#include <string.h>
#include <time.h>

typedef int MPI_Comm;
typedef int MPI_Datatype;
#define MPI_COMM_WORLD 0
#define MPI_FLOAT 4 /* sizeof */
#define MPI_INT 4
#define MPI_BYTE 1
typedef int MPI_Op;
#define MPI_MIN 0

static inline int MPI_Init(int *, char ***) { return 0; }
static inline int MPI_Finalize(void) { return 0; }
static inline int MPI_Comm_size(MPI_Comm, int *n) { *n = 1; return 0; }
static inline int MPI_Comm_rank(MPI_Comm, int *r) { *r = 0; return 0; }
static inline int MPI_Barrier(MPI_Comm) { return 0; }
static inline double MPI_Wtime(void) {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}
static inline int MPI_Gather(const void *send, int count, MPI_Datatype dt, void *recv, int, MPI_Datatype,
                             int, MPI_Comm) {
    if (recv && send) memcpy(recv, send, (size_t)count * (size_t)dt);
    return 0;
}
static inline int MPI_Bcast(void *, int, MPI_Datatype, int, MPI_Comm) { return 0; }
static inline int MPI_Allreduce(const void *send, void *recv, int count, MPI_Datatype dt, MPI_Op, MPI_Comm) {
    if (recv && send && recv != send) memcpy(recv, send, (size_t)count * (size_t)dt);
    return 0;
}
#endif
