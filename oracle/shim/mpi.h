/* Single-rank MPI shim — TEST INFRASTRUCTURE ONLY (oracle/, never shipped, never on the product path).
 *
 * The reference (rometsch/fargocpt) is an MPI+OpenMP program; this container has no MPI.
 * To build the UNMODIFIED reference sources as the parity oracle (oracle/_ref/fargocpt_exe)
 * we provide the subset of the MPI C API the reference calls, specialised to one rank:
 * collectives are memcpy, point-to-point is unreachable (abort), MPI-IO maps to stdio with
 * seek offsets scaled by the etype size given to MPI_File_set_view (every reference call
 * site uses etype MPI_DOUBLE: polargrid.cpp:148-151, radialgrid.cpp:186-192).
 */
#ifndef ORACLE_SHIM_MPI_H
#define ORACLE_SHIM_MPI_H
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef int MPI_Comm;
typedef int MPI_Datatype; /* value = size in bytes */
typedef int MPI_Op;
typedef int MPI_Info;
typedef int MPI_Request;
typedef long MPI_Aint;
typedef long long MPI_Offset;
typedef struct { int MPI_SOURCE, MPI_TAG, MPI_ERROR, count_bytes; } MPI_Status;
typedef struct shim_mpi_file { FILE *fp; int etype; } *MPI_File;

#define MPI_COMM_WORLD 0
#define MPI_INFO_NULL 0
#define MPI_SUCCESS 0
#define MPI_STATUS_IGNORE ((MPI_Status *)0)
#define MPI_CHAR 1
#define MPI_INT 4
#define MPI_UNSIGNED 4
#define MPI_INT32_T 4
#define MPI_UINT32_T 4
#define MPI_DOUBLE 8
#define MPI_AINT 8
#define MPI_UNSIGNED_LONG 8
#define MPI_SUM 1
#define MPI_MIN 2
#define MPI_MAX 3
#define MPI_THREAD_SINGLE 0
#define MPI_THREAD_FUNNELED 1
#define MPI_THREAD_SERIALIZED 2
#define MPI_THREAD_MULTIPLE 3
#define MPI_MODE_RDONLY 1
#define MPI_MODE_WRONLY 2
#define MPI_MODE_CREATE 4
#define MPI_MODE_APPEND 8
#define MPI_SEEK_SET 0
#define MPI_SEEK_END 2
#define MPI_MAX_PROCESSOR_NAME 256
#define MPI_MAX_ERROR_STRING 256

static inline void shim_unreachable(const char *what)
{
    fprintf(stderr, "mpi shim: %s is unreachable with a single rank\n", what);
    abort();
}

static inline int MPI_Init_thread(int *argc, char ***argv, int required, int *provided)
{ (void)argc; (void)argv; *provided = required; return MPI_SUCCESS; }
static inline int MPI_Initialized(int *flag) { *flag = 1; return MPI_SUCCESS; }
static inline int MPI_Finalize(void) { return MPI_SUCCESS; }
static inline int MPI_Abort(MPI_Comm c, int code) { (void)c; exit(code); return MPI_SUCCESS; }
static inline int MPI_Comm_rank(MPI_Comm c, int *rank) { (void)c; *rank = 0; return MPI_SUCCESS; }
static inline int MPI_Comm_size(MPI_Comm c, int *size) { (void)c; *size = 1; return MPI_SUCCESS; }
static inline int MPI_Barrier(MPI_Comm c) { (void)c; return MPI_SUCCESS; }
static inline int MPI_Get_processor_name(char *name, int *len)
{ strcpy(name, "localhost"); *len = 9; return MPI_SUCCESS; }
static inline int MPI_Error_class(int err, int *cls) { *cls = err; return MPI_SUCCESS; }
static inline int MPI_Error_string(int err, char *s, int *len)
{ *len = snprintf(s, MPI_MAX_ERROR_STRING, "mpi shim error %d", err); return MPI_SUCCESS; }

static inline int shim_copy(const void *s, void *r, int count, MPI_Datatype t)
{ if (s != r) memmove(r, s, (size_t)count * (size_t)t); return MPI_SUCCESS; }
static inline int MPI_Allreduce(const void *s, void *r, int n, MPI_Datatype t, MPI_Op op, MPI_Comm c)
{ (void)op; (void)c; return shim_copy(s, r, n, t); }
static inline int MPI_Reduce(const void *s, void *r, int n, MPI_Datatype t, MPI_Op op, int root, MPI_Comm c)
{ (void)op; (void)root; (void)c; return shim_copy(s, r, n, t); }
static inline int MPI_Bcast(void *b, int n, MPI_Datatype t, int root, MPI_Comm c)
{ (void)b; (void)n; (void)t; (void)root; (void)c; return MPI_SUCCESS; }
static inline int MPI_Gather(const void *s, int sn, MPI_Datatype st, void *r, int rn, MPI_Datatype rt, int root, MPI_Comm c)
{ (void)rn; (void)rt; (void)root; (void)c; return shim_copy(s, r, sn, st); }
static inline int MPI_Allgather(const void *s, int sn, MPI_Datatype st, void *r, int rn, MPI_Datatype rt, MPI_Comm c)
{ (void)rn; (void)rt; (void)c; return shim_copy(s, r, sn, st); }
static inline int MPI_Gatherv(const void *s, int sn, MPI_Datatype st, void *r, const int *rn, const int *displs,
			      MPI_Datatype rt, int root, MPI_Comm c)
{ (void)rn; (void)root; (void)c; return shim_copy(s, (char *)r + (size_t)displs[0] * (size_t)rt, sn, st); }

static inline int MPI_Send(const void *b, int n, MPI_Datatype t, int d, int tag, MPI_Comm c)
{ (void)b; (void)n; (void)t; (void)d; (void)tag; (void)c; shim_unreachable("MPI_Send"); return 1; }
static inline int MPI_Ssend(const void *b, int n, MPI_Datatype t, int d, int tag, MPI_Comm c)
{ (void)b; (void)n; (void)t; (void)d; (void)tag; (void)c; shim_unreachable("MPI_Ssend"); return 1; }
static inline int MPI_Recv(void *b, int n, MPI_Datatype t, int s, int tag, MPI_Comm c, MPI_Status *st)
{ (void)b; (void)n; (void)t; (void)s; (void)tag; (void)c; (void)st; shim_unreachable("MPI_Recv"); return 1; }
static inline int MPI_Isend(const void *b, int n, MPI_Datatype t, int d, int tag, MPI_Comm c, MPI_Request *r)
{ (void)b; (void)n; (void)t; (void)d; (void)tag; (void)c; (void)r; shim_unreachable("MPI_Isend"); return 1; }
static inline int MPI_Irecv(void *b, int n, MPI_Datatype t, int s, int tag, MPI_Comm c, MPI_Request *r)
{ (void)b; (void)n; (void)t; (void)s; (void)tag; (void)c; (void)r; shim_unreachable("MPI_Irecv"); return 1; }
static inline int MPI_Wait(MPI_Request *r, MPI_Status *st) { (void)r; (void)st; return MPI_SUCCESS; }
static inline int MPI_Probe(int s, int tag, MPI_Comm c, MPI_Status *st)
{ (void)s; (void)tag; (void)c; (void)st; shim_unreachable("MPI_Probe"); return 1; }
static inline int MPI_Get_count(const MPI_Status *st, MPI_Datatype t, int *n)
{ *n = st ? st->count_bytes / (t ? t : 1) : 0; return MPI_SUCCESS; }

static inline int MPI_Get_address(const void *p, MPI_Aint *a) { *a = (MPI_Aint)(intptr_t)p; return MPI_SUCCESS; }
static inline int MPI_Type_create_struct(int n, const int *lens, const MPI_Aint *offs, const MPI_Datatype *types, MPI_Datatype *out)
{
    long ext = 0;
    for (int i = 0; i < n; ++i) { long e = offs[i] + (long)lens[i] * types[i]; if (e > ext) ext = e; }
    *out = (MPI_Datatype)((ext + 7) / 8 * 8);
    return MPI_SUCCESS;
}
static inline int MPI_Type_indexed(int n, const int *lens, const int *offs, MPI_Datatype old, MPI_Datatype *out)
{ (void)n; (void)lens; (void)offs; *out = old; return MPI_SUCCESS; }
static inline int MPI_Type_commit(MPI_Datatype *t) { (void)t; return MPI_SUCCESS; }
static inline int MPI_Type_free(MPI_Datatype *t) { (void)t; return MPI_SUCCESS; }

static inline int MPI_File_open(MPI_Comm c, const char *name, int amode, MPI_Info info, MPI_File *fh)
{
    (void)c; (void)info;
    const char *mode = "rb";
    if (amode & MPI_MODE_WRONLY) {
	if (amode & MPI_MODE_APPEND) mode = "ab";
	else {
	    /* MPI_MODE_CREATE|WRONLY does not truncate: open r+b if the file exists, else create */
	    FILE *probe = fopen(name, "r+b");
	    if (probe) { *fh = (MPI_File)malloc(sizeof(**fh)); (*fh)->fp = probe; (*fh)->etype = 1; return MPI_SUCCESS; }
	    mode = "w+b";
	}
    }
    FILE *fp = fopen(name, mode);
    if (!fp) { *fh = 0; return 1; }
    *fh = (MPI_File)malloc(sizeof(**fh));
    (*fh)->fp = fp;
    (*fh)->etype = 1;
    return MPI_SUCCESS;
}
static inline int MPI_File_close(MPI_File *fh) { if (*fh) { fclose((*fh)->fp); free(*fh); *fh = 0; } return MPI_SUCCESS; }
static inline int MPI_File_set_view(MPI_File fh, MPI_Offset disp, MPI_Datatype etype, MPI_Datatype ftype, const char *rep, MPI_Info info)
{ (void)ftype; (void)rep; (void)info; fh->etype = etype; fseek(fh->fp, (long)disp, SEEK_SET); return MPI_SUCCESS; }
static inline int MPI_File_seek(MPI_File fh, MPI_Offset off, int whence)
{ fseek(fh->fp, (long)(off * fh->etype), whence == MPI_SEEK_END ? SEEK_END : SEEK_SET); return MPI_SUCCESS; }
static inline int MPI_File_get_size(MPI_File fh, MPI_Offset *size)
{ long cur = ftell(fh->fp); fseek(fh->fp, 0, SEEK_END); *size = ftell(fh->fp); fseek(fh->fp, cur, SEEK_SET); return MPI_SUCCESS; }
static inline int MPI_File_write(MPI_File fh, const void *buf, int n, MPI_Datatype t, MPI_Status *st)
{ (void)st; return fwrite(buf, (size_t)t, (size_t)n, fh->fp) == (size_t)n ? MPI_SUCCESS : 1; }
static inline int MPI_File_write_all(MPI_File fh, const void *buf, int n, MPI_Datatype t, MPI_Status *st)
{ return MPI_File_write(fh, buf, n, t, st); }
static inline int MPI_File_read_all(MPI_File fh, void *buf, int n, MPI_Datatype t, MPI_Status *st)
{ (void)st; return fread(buf, (size_t)t, (size_t)n, fh->fp) == (size_t)n ? MPI_SUCCESS : 1; }

#ifdef __cplusplus
}
#endif
#endif
