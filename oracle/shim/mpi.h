/* mpi.h -- single-host stand-in for the MPI subset used by the reference pgsd.c
 *
 * TEST INFRASTRUCTURE ONLY (oracle/).  Nothing in the product (pgsd_sph_b200/)
 * includes or links this.  It exists so that the UNMODIFIED reference source
 * /root/reference/pgsd/pgsd/pgsd.c can be compiled in place into oracle/_ref/
 * and run at P = 1..N "ranks" on one host without an MPI installation.
 *
 * Process model: MPI_Init() forks P-1 children (P from env PGSD_SHIM_NP);
 * collectives go through a MAP_SHARED scratch area guarded by a
 * PTHREAD_PROCESS_SHARED barrier; MPI_File_* map onto open/pread/pwrite.
 *
 * Surface covered = exactly what the reference uses:
 *   pgsd.c:106-202 (Bcast/Allreduce/Barrier helpers), :1126 (Allgather),
 *   :1015-1074, :1154, :1289-1306, :1456-1469, :1500-1520, :1748, :1788-1798,
 *   :1906, :2032, :2229, :2534 (MPI_File_*), pgsd.h:19-31 (my_MPI_SIZE_T).
 */
#ifndef PGSD_ORACLE_SHIM_MPI_H
#define PGSD_ORACLE_SHIM_MPI_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef int MPI_Comm;
typedef int MPI_Datatype;
typedef int MPI_Op;
typedef int MPI_Info;
typedef long long MPI_Offset; /* reference passes &(long long file_size): pgsd.c:1519 */
typedef long MPI_Aint;
typedef struct { int unused; } MPI_Status;

/* MPI_File must be a pointer type: the reference tests handle->fh == NULL
   (pgsd.c:1494,1766,1805). */
struct shim_file;
typedef struct shim_file* MPI_File;

#define MPI_COMM_WORLD 0
#define MPI_SUCCESS 0
#define MPI_ERR_OTHER 15
#define MPI_INFO_NULL 0
#define MPI_STATUS_IGNORE ((MPI_Status*)0)
#define MPI_IN_PLACE ((void*)1)

/* datatype codes: low byte = element size, next byte = kind (0 unsigned, 1 signed, 2 float) */
#define SHIM_DT(kind, size) (((kind) << 8) | (size))
#define MPI_BYTE SHIM_DT(0, 1)
#define MPI_UINT8_T SHIM_DT(0, 1)
#define MPI_UNSIGNED_CHAR SHIM_DT(0, 1)
#define MPI_UINT16_T SHIM_DT(0, 2)
#define MPI_UNSIGNED_SHORT SHIM_DT(0, 2)
#define MPI_UINT32_T SHIM_DT(0, 4)
#define MPI_UNSIGNED SHIM_DT(0, 4)
#define MPI_UINT64_T SHIM_DT(0, 8)
#define MPI_UNSIGNED_LONG SHIM_DT(0, 8)
#define MPI_UNSIGNED_LONG_LONG SHIM_DT(0, 8)
#define MPI_INT SHIM_DT(1, 4)
#define MPI_INT64_T SHIM_DT(1, 8)
#define MPI_LONG_LONG_INT SHIM_DT(1, 8)
#define MPI_DOUBLE SHIM_DT(2, 8)

#define MPI_SUM 1
#define MPI_MIN 2
#define MPI_MAX 3

#define MPI_MODE_RDONLY 2
#define MPI_MODE_RDWR 8
#define MPI_MODE_CREATE 1
#define MPI_MODE_EXCL 64

#define MPI_SEEK_SET 600
#define MPI_SEEK_END 604

int MPI_Init(int* argc, char*** argv);
int MPI_Finalize(void);
int MPI_Comm_rank(MPI_Comm comm, int* rank);
int MPI_Comm_size(MPI_Comm comm, int* size);
int MPI_Barrier(MPI_Comm comm);
int MPI_Bcast(void* buf, int count, MPI_Datatype dt, int root, MPI_Comm comm);
int MPI_Allreduce(const void* send, void* recv, int count, MPI_Datatype dt, MPI_Op op,
                  MPI_Comm comm);
int MPI_Allgather(const void* send, int scount, MPI_Datatype sdt, void* recv, int rcount,
                  MPI_Datatype rdt, MPI_Comm comm);

int MPI_Type_create_struct(int n, const int* blocklens, const MPI_Aint* displs,
                           const MPI_Datatype* types, MPI_Datatype* newtype);
int MPI_Type_commit(MPI_Datatype* dt);
int MPI_Type_free(MPI_Datatype* dt);

int MPI_File_open(MPI_Comm comm, const char* fname, int amode, MPI_Info info, MPI_File* fh);
int MPI_File_close(MPI_File* fh);
int MPI_File_set_size(MPI_File fh, MPI_Offset size);
int MPI_File_get_size(MPI_File fh, MPI_Offset* size);
int MPI_File_seek(MPI_File fh, MPI_Offset off, int whence);
int MPI_File_read(MPI_File fh, void* buf, int count, MPI_Datatype dt, MPI_Status* st);
int MPI_File_write(MPI_File fh, const void* buf, int count, MPI_Datatype dt, MPI_Status* st);
int MPI_File_read_at(MPI_File fh, MPI_Offset off, void* buf, int count, MPI_Datatype dt,
                     MPI_Status* st);
int MPI_File_write_at(MPI_File fh, MPI_Offset off, const void* buf, int count, MPI_Datatype dt,
                      MPI_Status* st);

/* shim extras used by the oracle driver */
double shim_wtime(void);
void* shim_shared_alloc(size_t bytes); /* MAP_SHARED|MAP_ANONYMOUS; call before MPI_Init */

#ifdef __cplusplus
}
#endif
#endif
