/* -*- C++ -*-
 * Single-process functional MPI shim (TEST INFRASTRUCTURE ONLY).
 *
 * The reference (amanotk/pic-nix) includes <mpi.h> unconditionally (nix/nix.hpp:31) and routes every
 * halo message -- even between chunks of the same rank -- through MPI point-to-point calls
 * (nix/chunk.hpp:464-543, nix/chunk.cpp:288-395).  This container has no MPI, so to compile and run
 * the UNMODIFIED reference sources as the parity oracle (oracle/_ref) we provide the few MPI entry
 * points those files touch, implemented as an in-process mailbox keyed by (communicator, tag).
 * Everything lives in rank 0 of a world of size 1.
 *
 * This file is not part of the product; only oracle/ref_driver.cpp is compiled against it.
 */
#ifndef PICNIX_ORACLE_MPI_SHIM_H
#define PICNIX_ORACLE_MPI_SHIM_H

#include <cstddef>
#include <cstdint>

typedef int MPI_Comm;
typedef int MPI_Request;
typedef int MPI_Datatype; /* value == size of one element in bytes */
typedef int MPI_Info;
typedef int MPI_Op;
typedef int MPI_File;
typedef long long MPI_Offset;
typedef long long MPI_Aint;

struct MPI_Status {
  int MPI_SOURCE;
  int MPI_TAG;
  int MPI_ERROR;
  int shim_bytes;
};

#define MPI_SUCCESS 0
#define MPI_COMM_WORLD 0
#define MPI_COMM_NULL (-1)
#define MPI_PROC_NULL (-2)
#define MPI_ANY_SOURCE (-3)
#define MPI_ANY_TAG (-4)
#define MPI_REQUEST_NULL (-1)
#define MPI_INFO_NULL 0
#define MPI_STATUS_IGNORE ((MPI_Status*)0)
#define MPI_STATUSES_IGNORE ((MPI_Status*)0)
#define MPI_IN_PLACE ((void*)1)

#define MPI_BYTE 1
#define MPI_CHAR 1
#define MPI_CXX_BOOL 1
#define MPI_INT 4
#define MPI_FLOAT 4
#define MPI_DOUBLE 8
#define MPI_INT64_T 8
#define MPI_LONG_LONG 8
/* derived datatypes (MPI_Type_contiguous / create_hindexed / create_subarray) get handles >= this */
#define PICNIX_SHIM_DERIVED_BASE 0x10000

#define MPI_SUM 1
#define MPI_LAND 2
#define MPI_MIN 3
#define MPI_MAX 4

#define MPI_THREAD_SINGLE 0
#define MPI_THREAD_FUNNELED 1
#define MPI_THREAD_SERIALIZED 2
#define MPI_THREAD_MULTIPLE 3

#define MPI_COMM_TYPE_SHARED 1
#define MPI_ORDER_C 0
#define MPI_ORDER_FORTRAN 1
#define MPI_MODE_RDONLY 1
#define MPI_MODE_WRONLY 2
#define MPI_MODE_CREATE 4
#define MPI_MODE_APPEND 8
#define MPI_MODE_RDWR 16
#define MPI_SEEK_SET 0
#define MPI_SEEK_CUR 1
#define MPI_SEEK_END 2

#ifdef __cplusplus
extern "C" {
#endif

int MPI_Init(int*, char***);
int MPI_Init_thread(int*, char***, int required, int* provided);
int MPI_Finalize(void);
int MPI_Abort(MPI_Comm, int code);
int MPI_Comm_rank(MPI_Comm, int* rank);
int MPI_Comm_size(MPI_Comm, int* size);
int MPI_Comm_dup(MPI_Comm, MPI_Comm* newcomm);
int MPI_Comm_free(MPI_Comm*);
int MPI_Comm_split(MPI_Comm, int color, int key, MPI_Comm* newcomm);
int MPI_Comm_split_type(MPI_Comm, int type, int key, MPI_Info, MPI_Comm* newcomm);
int MPI_Barrier(MPI_Comm);
int MPI_Bcast(void* buf, int count, MPI_Datatype, int root, MPI_Comm);
int MPI_Allreduce(const void* sbuf, void* rbuf, int count, MPI_Datatype, MPI_Op, MPI_Comm);
int MPI_Reduce(const void* sbuf, void* rbuf, int count, MPI_Datatype, MPI_Op, int root, MPI_Comm);
int MPI_Allgather(const void* sbuf, int scount, MPI_Datatype, void* rbuf, int rcount, MPI_Datatype, MPI_Comm);
int MPI_Gather(const void* sbuf, int scount, MPI_Datatype, void* rbuf, int rcount, MPI_Datatype, int root,
               MPI_Comm);
int MPI_Gatherv(const void* sbuf, int scount, MPI_Datatype, void* rbuf, const int* rcounts, const int* displs,
                MPI_Datatype, int root, MPI_Comm);
int MPI_Allgatherv(const void* sbuf, int scount, MPI_Datatype, void* rbuf, const int* rcounts,
                   const int* displs, MPI_Datatype, MPI_Comm);
int MPI_Isend(const void* buf, int count, MPI_Datatype, int dest, int tag, MPI_Comm, MPI_Request*);
int MPI_Irecv(void* buf, int count, MPI_Datatype, int source, int tag, MPI_Comm, MPI_Request*);
int MPI_Iprobe(int source, int tag, MPI_Comm, int* flag, MPI_Status*);
int MPI_Get_count(const MPI_Status*, MPI_Datatype, int* count);
int MPI_Type_size(MPI_Datatype, int* size);
int MPI_Wait(MPI_Request*, MPI_Status*);
int MPI_Waitall(int n, MPI_Request*, MPI_Status*);
int MPI_Testall(int n, MPI_Request*, int* flag, MPI_Status*);

/* derived datatypes: enough for nix/nixio.cpp (one contiguous block per process) */
int MPI_Type_contiguous(int count, MPI_Datatype oldtype, MPI_Datatype* newtype);
int MPI_Type_create_hindexed(int count, const int* blocklens, const MPI_Aint* displs, MPI_Datatype oldtype,
                             MPI_Datatype* newtype);
int MPI_Type_create_subarray(int ndim, const int* gshape, const int* lshape, const int* offset, int order,
                             MPI_Datatype oldtype, MPI_Datatype* newtype);
int MPI_Type_commit(MPI_Datatype*);
int MPI_Type_free(MPI_Datatype*);

/* MPI-IO on POSIX files, single process: the *_all calls write the process' block at
 * view displacement + block offset, the *_at calls at view displacement + offset * etype size */
int MPI_File_open(MPI_Comm, const char* filename, int amode, MPI_Info, MPI_File* fh);
int MPI_File_close(MPI_File* fh);
int MPI_File_delete(const char* filename, MPI_Info);
int MPI_File_seek(MPI_File fh, MPI_Offset offset, int whence);
int MPI_File_get_size(MPI_File fh, MPI_Offset* size);
int MPI_File_get_position(MPI_File fh, MPI_Offset* pos);
int MPI_File_set_view(MPI_File fh, MPI_Offset disp, MPI_Datatype etype, MPI_Datatype filetype,
                      const char* datarep, MPI_Info);
int MPI_File_iread_all(MPI_File fh, void* buf, int count, MPI_Datatype, MPI_Request*);
int MPI_File_iwrite_all(MPI_File fh, const void* buf, int count, MPI_Datatype, MPI_Request*);
int MPI_File_iread_at(MPI_File fh, MPI_Offset offset, void* buf, int count, MPI_Datatype, MPI_Request*);
int MPI_File_iwrite_at(MPI_File fh, MPI_Offset offset, const void* buf, int count, MPI_Datatype, MPI_Request*);

/* shim-only helper: drop every undelivered message and pending request */
void picnix_mpi_shim_reset(void);

#ifdef __cplusplus
}
#endif

#endif
