// comm.h -- rank communicator of libpgsd_b200 (replaces MPI_COMM_WORLD of the reference,
// /root/reference/pgsd/pgsd/pgsd.c:1487-1488 and the 70 collective call sites listed in
// SURVEY.md section 5.8).  The only primitive the file layer needs is an all-gather of
// small uint64 vectors; barrier and broadcast are built on it.
#pragma once
#include <cstddef>
#include <cstdint>
#include <string>

namespace pgsdb
{
enum class CommKind
    {
    Single,
    Host,
    Shm,
    Nccl
    };

struct SizeScan // K2 result for one chunk
    {
    uint64_t excl;  // sum over ranks < rank
    uint64_t total; // sum over all ranks
    uint64_t maxv;  // max over all ranks
    uint64_t first; // value of rank 0 (the index points at rank 0's copy of buffered chunks)
    };

class Comm
    {
    public:
    virtual ~Comm() { }
    int rank = 0;
    int nprocs = 1;
    CommKind kind = CommKind::Single;
    // recv[r * n + i] = value i of rank r.  0 on success.
    virtual int allgather(const uint64_t* send, uint64_t* recv, size_t n) = 0;
    // all-gather `n` sizes and reduce them per entry (K2).  The NCCL transport keeps the
    // gathered matrix on the device and runs the scan kernel there; host transports use
    // scan_sizes_host().
    virtual int allgather_scan(const uint64_t* send, SizeScan* out, size_t n);
    int barrier();
    };

// host reduction of a gathered [P][n] matrix (used by the host/shm transports)
void scan_sizes_host(const uint64_t* gathered, int P, size_t n, int rank, SizeScan* out);

Comm* comm();                 // never NULL (Single by default)
// Takes ownership; NULL -> back to Single.  Refused (false, `c` deleted) while file handles opened under the current
// communicator are still open: they keep a pointer to it.
bool comm_replace(Comm* c);
void comm_acquire(); // a file handle starts / stops using comm()
void comm_release();
int comm_users();
const char* comm_kind_name();

Comm* make_host_comm(int rank, int nprocs,
                     int (*fn)(void*, const uint64_t*, uint64_t*, size_t), void* ctx);
Comm* make_shm_comm(int rank, int nprocs, const char* name, std::string& err);
Comm* make_nccl_comm(int rank, int nprocs, const void* unique_id, int device, std::string& err);
int nccl_unique_id(void* out128, std::string& err);

extern uint64_t g_collectives; // statistics
} // namespace pgsdb
