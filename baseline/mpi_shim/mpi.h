/* Single-node stand-in for <mpi.h>, only so the UNMODIFIED reference programs (mpi/mpi_conv.c,
 * mpi_omp/mpi_omp_conv.c) can be compiled and run on a box that has no MPI installation.
 *
 * It implements exactly the subset those programs call: world communicator, one
 * process per rank started by baseline/mpi_shim/mpirun_shim, point-to-point messages over one
 * AF_UNIX socket pair per rank pair (FIFO, so MPI's non-overtaking rule holds), vector/contiguous
 * datatypes, Bcast/Barrier built on the same channels, MPI_File_* mapped to POSIX open/lseek/read/
 * write, MPI_Wtime on CLOCK_MONOTONIC. Nothing here is part of the B200 framework; it exists for the
 * reference arm of the benchmark and the same-box CPU comparison.
 */
#ifndef PCONV_MPI_SHIM_H
#define PCONV_MPI_SHIM_H

#ifdef __cplusplus
extern "C" {
#endif

typedef int MPI_Comm;
typedef int MPI_Datatype;
typedef int MPI_Info;
typedef long long MPI_Offset;
typedef struct { int MPI_SOURCE, MPI_TAG, MPI_ERROR, count_; } MPI_Status;
typedef struct pconv_mpi_request* MPI_Request;
typedef struct pconv_mpi_file* MPI_File;

#define MPI_SUCCESS 0
#define MPI_COMM_WORLD 0
#define MPI_INFO_NULL 0
#define MPI_REQUEST_NULL ((MPI_Request)0)
#define MPI_STATUS_IGNORE ((MPI_Status*)0)
#define MPI_PROC_NULL (-2)

#define MPI_BYTE 1
#define MPI_CHAR 2
#define MPI_INT 3
#define MPI_FLOAT 4
#define MPI_DOUBLE 5

#define MPI_MODE_RDONLY 2
#define MPI_MODE_RDWR 8
#define MPI_MODE_WRONLY 4
#define MPI_MODE_CREATE 1
#define MPI_SEEK_SET 600
#define MPI_SEEK_CUR 602
#define MPI_SEEK_END 604

int MPI_Init(int* argc, char*** argv);
int MPI_Finalize(void);
int MPI_Abort(MPI_Comm comm, int errorcode);
int MPI_Comm_size(MPI_Comm comm, int* size);
int MPI_Comm_rank(MPI_Comm comm, int* rank);
double MPI_Wtime(void);
int MPI_Barrier(MPI_Comm comm);
int MPI_Bcast(void* buf, int count, MPI_Datatype type, int root, MPI_Comm comm);

int MPI_Type_vector(int count, int blocklength, int stride, MPI_Datatype oldtype, MPI_Datatype* newtype);
int MPI_Type_contiguous(int count, MPI_Datatype oldtype, MPI_Datatype* newtype);
int MPI_Type_commit(MPI_Datatype* type);
int MPI_Type_free(MPI_Datatype* type);

int MPI_Send(const void* buf, int count, MPI_Datatype type, int dest, int tag, MPI_Comm comm);
int MPI_Recv(void* buf, int count, MPI_Datatype type, int source, int tag, MPI_Comm comm, MPI_Status* status);
int MPI_Isend(const void* buf, int count, MPI_Datatype type, int dest, int tag, MPI_Comm comm, MPI_Request* req);
int MPI_Irecv(void* buf, int count, MPI_Datatype type, int source, int tag, MPI_Comm comm, MPI_Request* req);
int MPI_Wait(MPI_Request* req, MPI_Status* status);

int MPI_File_open(MPI_Comm comm, const char* filename, int amode, MPI_Info info, MPI_File* fh);
int MPI_File_close(MPI_File* fh);
int MPI_File_seek(MPI_File fh, MPI_Offset offset, int whence);
int MPI_File_read(MPI_File fh, void* buf, int count, MPI_Datatype type, MPI_Status* status);
int MPI_File_write(MPI_File fh, const void* buf, int count, MPI_Datatype type, MPI_Status* status);

#ifdef __cplusplus
}
#endif
#endif
