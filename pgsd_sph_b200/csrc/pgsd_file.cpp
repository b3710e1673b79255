// pgsd_file.cpp -- the pgsd.h C ABI of libpgsd_b200: file create/open/close, chunk write,
// frame commit, flush, chunk lookup/read, index + namelist management.
//
// Written from scratch against the behaviour of the reference libpgsd
// (/root/reference/pgsd/pgsd/pgsd.c, cited per function below); files it writes are
// byte-identical to the reference's for the same call sequence and rank partitioning.
//
// Design differences (B200-first):
//   * Every rank keeps a replica of the namelist and index and runs the same layout state
//     machine, so the ~90-120 MPI collectives per frame of the reference
//     collapse into ONE all-gather of the frame's chunk-size vector + exclusive scan (K2) at
//     frame commit.  With more than one rank, pgsd_write_chunk only records the chunk (device
//     data is packed into the frame arena by K1; host data is copied into a staging buffer);
//     the layout is computed when the frame is committed (pgsd_end_frame / pgsd_flush /
//     pgsd_close / any lookup on a writable file).
//   * Device-resident chunk bytes go arena -> pinned ring -> pwrite on writer threads (K3) and
//     overlap the packing of the next frame.  Only rank 0 writes header / index / namelist.
//   * No MPI: ranks come from the communicator installed through pgsd_b200_comm_init_*().
#include "../../include/pgsd.h"
#include "../../include/pgsd_b200.h"
#include "device.h"
#include "file_internal.h"

#include <cerrno>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fcntl.h>
#include <string>
#include <sys/stat.h>
#include <sys/types.h>
#include <unistd.h>
#include <unordered_map>
#include <vector>

namespace pgsdb
{
namespace
    {
const uint64_t MAGIC_ID = 0x65DF65DF65DF65DFull; // ref: pgsd.c:54
enum : size_t
    {
    INITIAL_INDEX_SIZE = 128,         // ref: pgsd.c:57-60
    INITIAL_NAME_BUFFER_SIZE = 1024,  // ref: pgsd.c:63-66
    INITIAL_FRAME_INDEX_SIZE = 16,    // ref: pgsd.c:69-72
    INITIAL_WRITE_BUFFER_SIZE = 1024, // ref: pgsd.c:75-78
    COPY_ENTRIES = 256 * 1024
    };
const uint64_t DEFAULT_MAXIMUM_WRITE_BUFFER_SIZE = 64ull * 1024 * 1024; // ref: pgsd.c:81-84
const uint64_t DEFAULT_INDEX_ENTRIES_TO_BUFFER = 256ull * 1024;         // ref: pgsd.c:87-90

static_assert(sizeof(pgsd_header) == 256, "header layout");
static_assert(sizeof(pgsd_index_entry) == 32, "index entry layout");
static_assert(sizeof(pgsd_handle) == 544, "handle layout (ref: pgsd.h:297-353, LP64)");

inline uint32_t make_version(unsigned major, unsigned minor) { return (major << 16) | minor; }

// ------------------------------------------------------------------------------ raw file I/O
bool pwrite_all(int fd, const void* buf, uint64_t n, uint64_t off)
    {
    const char* p = (const char*)buf;
    while (n > 0)
        {
        ssize_t k = pwrite(fd, p, n, (off_t)off);
        if (k < 0)
            {
            if (errno == EINTR)
                continue;
            return false;
            }
        p += k;
        off += (uint64_t)k;
        n -= (uint64_t)k;
        }
    return true;
    }

// reads up to n bytes; returns the number read (short at end of file), -1 on error
// The file layer's own small writes (index entries, write buffers, names, header).  While device frames are being
// staged they are queued behind them on the staging pipeline's serial writer (one thread, submission order: see
// device.cu) instead of contending with it for the inode lock from the caller's thread; otherwise written at once.
// Everything queued is in the file after drain_all() (pgsd_flush, pgsd_close, reads, index relocation).
bool put_small(int fd, const void* buf, uint64_t n, uint64_t off)
    {
    if (dev_async_host_write(fd, buf, n, off))
        return true;
    return pwrite_all(fd, buf, n, off);
    }

int64_t pread_some(int fd, void* buf, uint64_t n, uint64_t off)
    {
    char* p = (char*)buf;
    uint64_t got = 0;
    while (got < n)
        {
        ssize_t k = pread(fd, p + got, n - got, (off_t)(off + got));
        if (k < 0)
            {
            if (errno == EINTR)
                continue;
            return -1;
            }
        if (k == 0)
            break;
        got += (uint64_t)k;
        }
    return (int64_t)got;
    }

// ------------------------------------------------------------------------------ buffers
// Growth rules are observable (the namelist is written with its reserved size), so they follow
// the reference: byte buffers double until size+n < reserved (ref: pgsd.c:490-525), index
// buffers double when full (ref: pgsd.c:764-797).
int bytes_allocate(pgsd_byte_buffer& b, size_t reserve)
    {
    if (b.data || reserve == 0 || b.reserved != 0 || b.size != 0)
        return PGSD_ERROR_INVALID_ARGUMENT;
    b.data = (char*)calloc(reserve, 1);
    if (!b.data)
        return PGSD_ERROR_MEMORY_ALLOCATION_FAILED;
    b.reserved = reserve;
    return PGSD_SUCCESS;
    }

int bytes_append(pgsd_byte_buffer& b, const void* data, size_t n)
    {
    if (b.data == nullptr || n == 0 || b.reserved == 0)
        return PGSD_ERROR_INVALID_ARGUMENT;
    if (b.size + n > b.reserved)
        {
        size_t cap = b.reserved * 2;
        while (b.size + n >= cap)
            cap *= 2;
        char* p = (char*)realloc(b.data, cap);
        if (!p)
            return PGSD_ERROR_MEMORY_ALLOCATION_FAILED;
        b.data = p;
        memset(b.data + b.size + n, 0, cap - (b.size + n));
        b.reserved = cap;
        }
    memcpy(b.data + b.size, data, n);
    b.size += n;
    return PGSD_SUCCESS;
    }

void bytes_free(pgsd_byte_buffer& b)
    {
    free(b.data);
    b.data = nullptr;
    b.size = b.reserved = 0;
    }

int index_allocate(pgsd_index_buffer& b, size_t reserve)
    {
    if (b.data || reserve == 0 || b.reserved != 0 || b.size != 0)
        return PGSD_ERROR_INVALID_ARGUMENT;
    b.data = (pgsd_index_entry*)calloc(reserve, sizeof(pgsd_index_entry));
    if (!b.data)
        return PGSD_ERROR_MEMORY_ALLOCATION_FAILED;
    b.reserved = reserve;
    b.mapped_data = nullptr;
    b.mapped_len = 0;
    return PGSD_SUCCESS;
    }

int index_add(pgsd_index_buffer& b, const pgsd_index_entry& e)
    {
    if (b.reserved == 0)
        return PGSD_ERROR_INVALID_ARGUMENT;
    if (b.size == b.reserved)
        {
        size_t cap = b.reserved * 2;
        pgsd_index_entry* p = (pgsd_index_entry*)realloc(b.data, cap * sizeof(pgsd_index_entry));
        if (!p)
            return PGSD_ERROR_MEMORY_ALLOCATION_FAILED;
        b.data = p;
        memset(b.data + b.reserved, 0, (cap - b.reserved) * sizeof(pgsd_index_entry));
        b.reserved = cap;
        }
    b.data[b.size++] = e;
    return PGSD_SUCCESS;
    }

void index_free(pgsd_index_buffer& b)
    {
    free(b.data);
    memset(&b, 0, sizeof(b));
    }

// order of the on-disk index: by frame, then by name id (ref: pgsd.c:799-833)
inline int entry_cmp(const pgsd_index_entry& a, const pgsd_index_entry& b)
    {
    if (a.frame != b.frame)
        return a.frame < b.frame ? -1 : 1;
    if (a.id != b.id)
        return a.id < b.id ? -1 : 1;
    return 0;
    }

// In-place binary max-heap sort.  Not stable, and duplicate (frame, id) keys do occur (the same
// name written twice in one frame), so the exact procedure is observable in the file: sift-down
// heapify from the last parent, then repeatedly swap the root behind the shrinking heap
// (ref: pgsd.c:836-953).
void sift_down(pgsd_index_entry* v, size_t root, size_t last)
    {
    for (;;)
        {
        size_t child = 2 * root + 1;
        if (child > last)
            return;
        size_t pick = root;
        if (entry_cmp(v[pick], v[child]) < 0)
            pick = child;
        if (child + 1 <= last && entry_cmp(v[pick], v[child + 1]) < 0)
            pick = child + 1;
        if (pick == root)
            return;
        pgsd_index_entry t = v[root];
        v[root] = v[pick];
        v[pick] = t;
        root = pick;
        }
    }

void index_sort(pgsd_index_buffer& b)
    {
    if (b.size <= 1)
        return;
    pgsd_index_entry* v = b.data;
    for (size_t s = (b.size - 2) / 2 + 1; s-- > 0;)
        sift_down(v, s, b.size - 1);
    for (size_t end = b.size - 1; end > 0;)
        {
        pgsd_index_entry t = v[end];
        v[end] = v[0];
        v[0] = t;
        end--;
        sift_down(v, 0, end);
        }
    }

// ------------------------------------------------------------------------------ per-file state
enum class Src
    {
    HostUser,  // caller's pointer, valid only while its pgsd_write_chunk call runs (1 rank)
    HostStage, // copy in FileState::stage (deferred, > 1 rank)
    Device     // packed chunk in the current frame arena
    };

struct PendingOp // one recorded pgsd_write_chunk
    {
    pgsd_index_entry entry; // frame, id, type, N, M filled; location decided at commit
    uint64_t size;          // this rank's bytes
    uint64_t offset_bytes;  // this rank's start inside a direct chunk; UINT64_MAX = auto (K2 prefix)
    bool n_global_auto;
    bool all;
    Src src;
    const void* ptr;    // HostUser / Device
    uint64_t stage_off; // HostStage
    };

struct FileState
    {
    int fd = -1;
    Comm* comm = nullptr;
    std::unordered_map<std::string, uint16_t> name_ids;
    std::vector<PendingOp> ops;
    std::vector<char> stage;
    std::vector<WriteJob> dev_jobs; // device chunks of the frame being assembled
    bool device_frame_open = false;
    void* dev_frame = nullptr;      // this handle's frame arena (device.cu) while a frame is being assembled
    // rank-replicated view of the per-rank write buffers (ref: each rank's write_buffer.size)
    uint64_t root_wb_size = 0; // rank 0's write_buffer.size  (index locations refer to it)
    uint64_t wb_excl = 0;      // sum over ranks < rank of their write_buffer.size
    uint64_t wb_total = 0;     // sum over all ranks
    size_t file_index_cap = 0; // allocated entries behind handle->file_index.data
    };

inline FileState* state_of(pgsd_handle* h) { return (FileState*)h->fh; }

bool is_version2(const pgsd_handle* h) { return h->header.pgsd_version >= make_version(2, 0); }

// ref: pgsd.c:414-450
bool entry_valid(const pgsd_handle* h, size_t idx)
    {
    const pgsd_index_entry e = h->file_index.data[idx];
    const size_t es = type_size(e.type);
    if (es == 0)
        return false;
    uint64_t size = e.N * e.M * es;
    if ((uint64_t)e.location + size > (uint64_t)h->file_size)
        return false;
    if (e.frame >= h->header.index_allocated_entries)
        return false;
    if (e.id >= h->file_names.n_names + h->frame_names.n_names)
        return false;
    if (e.flags != 0)
        return false;
    return true;
    }

// Number of used entries of the index block mirrored in h->file_index.data[0..reserved):
// the first entry with location == 0 ends the list (ref: pgsd.c:660-705).
int index_count_used(pgsd_handle* h)
    {
    pgsd_index_buffer& b = h->file_index;
    if (b.data[0].location != 0 && !entry_valid(h, 0))
        return PGSD_ERROR_FILE_CORRUPT;
    if (b.data[0].location == 0)
        {
        b.size = 0;
        return PGSD_SUCCESS;
        }
    size_t L = 0, R = b.reserved;
    do
        {
        size_t m = (L + R) / 2;
        if (b.data[m].location != 0 && (!entry_valid(h, m) || b.data[m].frame < b.data[L].frame))
            return PGSD_ERROR_FILE_CORRUPT;
        if (b.data[m].location != 0)
            L = m;
        else
            R = m;
        } while (R - L > 1);
    b.size = R;
    return PGSD_SUCCESS;
    }

// ref: pgsd.c:602-707 (read variant; every rank loads its replica)
int index_load(pgsd_handle* h)
    {
    FileState* s = state_of(h);
    pgsd_index_buffer& b = h->file_index;
    const uint64_t n = h->header.index_allocated_entries;
    // divide instead of multiply: a corrupt header must not overflow the check
    if (n > ((uint64_t)h->file_size) / sizeof(pgsd_index_entry)
        || h->header.index_location + sizeof(pgsd_index_entry) * n > (uint64_t)h->file_size)
        return PGSD_ERROR_FILE_CORRUPT;
    if (n == 0)
        return PGSD_ERROR_INVALID_ARGUMENT;
    int rc = index_allocate(b, n);
    if (rc != PGSD_SUCCESS)
        return rc;
    s->file_index_cap = n;
    int64_t got = pread_some(s->fd, b.data, sizeof(pgsd_index_entry) * n, h->header.index_location);
    if (got != (int64_t)(sizeof(pgsd_index_entry) * n))
        return PGSD_ERROR_IO;
    dev_stats().file_bytes_read += (uint64_t)got;
    return index_count_used(h);
    }

void release_state(pgsd_handle* h)
    {
    FileState* s = state_of(h);
    index_free(h->file_index);
    index_free(h->frame_index);
    index_free(h->buffer_index);
    bytes_free(h->write_buffer);
    bytes_free(h->file_names.data);
    bytes_free(h->frame_names.data);
    h->file_names.n_names = 0;
    h->frame_names.n_names = 0;
    if (s)
        {
        if (s->dev_frame)
            dev_frame_abandon(s->dev_frame); // chunks packed for a frame that was never committed
        if (s->fd >= 0)
            close(s->fd);
        if (s->comm)
            comm_release();
        delete s;
        }
    h->fh = nullptr;
    }

// every rank learns whether all ranks are fine; returns the first non-zero code
int agree(Comm* c, int rc)
    {
    if (c->nprocs == 1)
        return rc;
    uint64_t mine = (uint64_t)(int64_t)rc;
    std::vector<uint64_t> all((size_t)c->nprocs);
    if (c->allgather(&mine, all.data(), 1) != 0)
        return PGSD_ERROR_IO;
    for (uint64_t v : all)
        if ((int64_t)v != 0)
            return (int)(int64_t)v;
    return 0;
    }

// ref: pgsd.c:1484-1703
int initialize_handle(pgsd_handle* h)
    {
    FileState* s = state_of(h);
    h->rank = s->comm->rank;
    h->nprocs = s->comm->nprocs;

    int64_t got = pread_some(s->fd, &h->header, sizeof(pgsd_header), 0);
    if (got < 0)
        return PGSD_ERROR_IO;
    if (got < (int64_t)sizeof(pgsd_header))
        memset((char*)&h->header + got, 0, sizeof(pgsd_header) - (size_t)got);
    if (h->header.magic != MAGIC_ID)
        return PGSD_ERROR_NOT_A_PGSD_FILE;
    if (h->header.pgsd_version < make_version(1, 0) && h->header.pgsd_version != make_version(0, 3))
        return PGSD_ERROR_INVALID_PGSD_FILE_VERSION;
    if (h->header.pgsd_version >= make_version(3, 0))
        return PGSD_ERROR_INVALID_PGSD_FILE_VERSION;

    struct stat st;
    if (fstat(s->fd, &st) != 0)
        return PGSD_ERROR_IO;
    h->file_size = (long long)st.st_size;

    const uint64_t nl_entries = h->header.namelist_allocated_entries;
    if (nl_entries > (uint64_t)h->file_size / PGSD_NAME_SIZE
        || h->header.namelist_location + PGSD_NAME_SIZE * nl_entries > (uint64_t)h->file_size)
        return PGSD_ERROR_FILE_CORRUPT;

    // namelist block -> name/id map (first-seen order = id), used bytes
    const size_t nl_bytes = PGSD_NAME_SIZE * nl_entries;
    int rc = bytes_allocate(h->file_names.data, nl_bytes);
    if (rc != PGSD_SUCCESS)
        return rc;
    got = pread_some(s->fd, h->file_names.data.data, nl_bytes, h->header.namelist_location);
    if (got != (int64_t)nl_bytes)
        return PGSD_ERROR_IO;
    dev_stats().file_bytes_read += (uint64_t)got;
    pgsd_byte_buffer& nb = h->file_names.data;
    if (nb.data[nb.reserved - 1] != 0)
        return PGSD_ERROR_FILE_CORRUPT;
    size_t start = 0;
    h->file_names.n_names = 0;
    while (start < nb.reserved)
        {
        const char* name = nb.data + start;
        if (name[0] == 0)
            break;
        // first occurrence wins, as in the reference's chained hash lookup (pgsd.c:374-405)
        s->name_ids.emplace(std::string(name, strnlen(name, nb.reserved - start)),
                            (uint16_t)h->file_names.n_names);
        h->file_names.n_names++;
        if (!is_version2(h))
            start += PGSD_NAME_SIZE;
        else
            start += strnlen(name, nb.reserved - start) + 1;
        }
    nb.size = start;

    rc = index_load(h);
    if (rc != PGSD_SUCCESS)
        return rc;
    h->cur_frame = h->file_index.size == 0 ? 0 : h->file_index.data[h->file_index.size - 1].frame + 1;

    if (h->open_flags != PGSD_OPEN_READONLY)
        {
        rc = index_allocate(h->frame_index, INITIAL_FRAME_INDEX_SIZE);
        if (rc == PGSD_SUCCESS)
            rc = index_allocate(h->buffer_index, INITIAL_FRAME_INDEX_SIZE);
        if (rc == PGSD_SUCCESS)
            rc = bytes_allocate(h->write_buffer, INITIAL_WRITE_BUFFER_SIZE);
        if (rc == PGSD_SUCCESS)
            rc = bytes_allocate(h->frame_names.data, PGSD_NAME_SIZE);
        if (rc != PGSD_SUCCESS)
            return rc;
        h->frame_names.n_names = 0;
        }
    h->pending_index_entries = 0;
    h->maximum_write_buffer_size = DEFAULT_MAXIMUM_WRITE_BUFFER_SIZE;
    h->index_entries_to_buffer = DEFAULT_INDEX_ENTRIES_TO_BUFFER;
    return PGSD_SUCCESS;
    }

// ------------------------------------------------------------------------------ layout state machine
// ref: pgsd.c:1216-1319
int flush_name_buffer(pgsd_handle* h)
    {
    FileState* s = state_of(h);
    if (h->frame_names.n_names == 0)
        return PGSD_SUCCESS;
    if (h->frame_names.data.size == 0)
        return PGSD_ERROR_INVALID_ARGUMENT;
    const size_t old_reserved = h->file_names.data.reserved;
    const size_t old_size = h->file_names.data.size;
    int rc = bytes_append(h->file_names.data, h->frame_names.data.data, h->frame_names.data.size);
    if (rc != PGSD_SUCCESS)
        return rc;
    h->file_names.n_names += h->frame_names.n_names;
    h->frame_names.n_names = 0;
    h->frame_names.data.size = 0;
    memset(h->frame_names.data.data, 0, h->frame_names.data.reserved);

    pgsd_byte_buffer& nb = h->file_names.data;
    if (nb.reserved % PGSD_NAME_SIZE != 0)
        return PGSD_ERROR_INVALID_ARGUMENT;
    bool ok = true;
    if (nb.reserved > old_reserved)
        {
        // capacity doubled: the whole list moves to the end of the file, header is rewritten
        const uint64_t off = (uint64_t)h->file_size;
        if (h->rank == 0)
            ok = put_small(s->fd, nb.data, nb.reserved, off);
        h->file_size += (long long)nb.reserved;
        h->header.namelist_location = off;
        h->header.namelist_allocated_entries = nb.reserved / PGSD_NAME_SIZE;
        if (h->rank == 0)
            {
            ok = ok && put_small(s->fd, &h->header, sizeof(pgsd_header), 0);
            dev_stats().file_bytes_written += nb.reserved + sizeof(pgsd_header);
            }
        }
    else if (h->rank == 0)
        {
        // in place: everything from the old end of the list to the end of the block
        ok = put_small(s->fd, nb.data + old_size, nb.reserved - old_size,
                        h->header.namelist_location + old_size);
        dev_stats().file_bytes_written += nb.reserved - old_size;
        }
    return ok ? PGSD_SUCCESS : PGSD_ERROR_IO;
    }

// ref: pgsd.c:1108-1201.  The reference all-gathers the per-rank buffer sizes here; the replica
// already holds their prefix/sum (accumulated from the K2 scan of each buffered chunk).
int flush_write_buffer(pgsd_handle* h)
    {
    FileState* s = state_of(h);
    if (s->root_wb_size == 0 && h->buffer_index.size == 0)
        return PGSD_SUCCESS;
    if (s->root_wb_size > 0 && h->buffer_index.size == 0)
        return PGSD_ERROR_INVALID_ARGUMENT;
    const uint64_t base = (uint64_t)h->file_size;
    bool ok = true;
    if (h->write_buffer.size > 0)
        {
        ok = put_small(s->fd, h->write_buffer.data, h->write_buffer.size, base + s->wb_excl);
        dev_stats().file_bytes_written += h->write_buffer.size;
        }
    h->file_size += (long long)s->wb_total;
    h->write_buffer.size = 0;
    s->root_wb_size = s->wb_excl = s->wb_total = 0;
    // the index points at rank 0's copy, which starts at `base`
    for (size_t i = 0; i < h->buffer_index.size; i++)
        {
        pgsd_index_entry e = h->buffer_index.data[i];
        e.location += (int64_t)base;
        int rc = index_add(h->frame_index, e);
        if (rc != PGSD_SUCCESS)
            return rc;
        }
    h->buffer_index.size = 0;
    return ok ? PGSD_SUCCESS : PGSD_ERROR_IO;
    }

// Submit the device chunks recorded so far to the staging pipeline (K3).
int submit_device_jobs(FileState* s)
    {
    if (!s->device_frame_open)
        return PGSD_SUCCESS;
    int rc = dev_frame_submit(s->fd, s->dev_jobs.data(), (int)s->dev_jobs.size(), s->dev_frame);
    if (rc != 0)
        dev_frame_abandon(s->dev_frame);
    s->dev_frame = nullptr;
    s->dev_jobs.clear();
    s->device_frame_open = false;
    return rc;
    }

// wait until every byte any rank has queued is in the file
int drain_all(pgsd_handle* h)
    {
    FileState* s = state_of(h);
    int rc = submit_device_jobs(s);
    if (rc == PGSD_SUCCESS)
        rc = dev_drain();
    return agree(s->comm, rc);
    }

// ref: pgsd.c:965-1091
int expand_file_index(pgsd_handle* h, size_t size_required)
    {
    FileState* s = state_of(h);
    if (h->open_flags == PGSD_OPEN_READONLY)
        return PGSD_ERROR_FILE_MUST_BE_WRITABLE;
    const size_t size_old = h->header.index_allocated_entries;
    size_t size_new = size_old * 2;
    while (size_new <= size_required)
        size_new *= 2;

    // The new block goes to the PHYSICAL end of the file (the reference asks MPI_File_get_size,
    // not handle->file_size, pgsd.c:1015), so every queued byte of every rank must have landed.
    int rc = drain_all(h);
    if (rc != PGSD_SUCCESS)
        return rc;
    uint64_t phys = 0;
    if (h->rank == 0)
        {
        struct stat st;
        if (fstat(s->fd, &st) != 0)
            return PGSD_ERROR_IO;
        phys = (uint64_t)st.st_size;
        }
    if (s->comm->nprocs > 1)
        {
        std::vector<uint64_t> all((size_t)s->comm->nprocs);
        if (s->comm->allgather(&phys, all.data(), 1) != 0)
            return PGSD_ERROR_IO;
        phys = all[0];
        }

    // grow the mirror: old block (stale tail included) followed by zeros
    if (s->file_index_cap < size_new)
        {
        pgsd_index_entry* p = (pgsd_index_entry*)realloc(h->file_index.data, size_new * sizeof(pgsd_index_entry));
        if (!p)
            return PGSD_ERROR_MEMORY_ALLOCATION_FAILED;
        h->file_index.data = p;
        s->file_index_cap = size_new;
        }
    memset(h->file_index.data + size_old, 0, (size_new - size_old) * sizeof(pgsd_index_entry));
    bool ok = true;
    if (h->rank == 0)
        {
        ok = pwrite_all(s->fd, h->file_index.data, size_new * sizeof(pgsd_index_entry), phys);
        dev_stats().file_bytes_written += size_new * sizeof(pgsd_index_entry);
        }
    h->header.index_location = phys;
    h->file_size = (long long)(phys + size_new * sizeof(pgsd_index_entry));
    h->header.index_allocated_entries = size_new;
    if (h->rank == 0)
        {
        ok = ok && put_small(s->fd, &h->header, sizeof(pgsd_header), 0);
        dev_stats().file_bytes_written += sizeof(pgsd_header);
        }
    h->file_index.reserved = size_new;
    // the reference re-reads the block and re-counts it (pgsd.c:1082-1088): entries of an
    // unfinished frame that an earlier mid-frame flush left behind the list become part of it
    rc = index_count_used(h);
    if (rc != PGSD_SUCCESS)
        return rc;
    return ok ? PGSD_SUCCESS : PGSD_ERROR_IO;
    }

// ref: pgsd.c:1955-2070
int flush_replica(pgsd_handle* h)
    {
    FileState* s = state_of(h);
    int rc = flush_name_buffer(h);
    if (rc != PGSD_SUCCESS)
        return rc;
    rc = flush_write_buffer(h);
    if (rc != PGSD_SUCCESS)
        return rc;
    if (h->pending_index_entries > h->frame_index.size)
        return PGSD_ERROR_INVALID_ARGUMENT;
    const uint64_t n_commit = h->frame_index.size - h->pending_index_entries;
    if (n_commit == 0)
        return PGSD_SUCCESS;
    if (h->file_index.size + n_commit > h->file_index.reserved)
        {
        rc = expand_file_index(h, h->file_index.size + n_commit);
        if (rc != PGSD_SUCCESS)
            return rc;
        }
    index_sort(h->frame_index);
    // The reference writes ALL frame_index entries -- the committed ones and, behind them, those
    // of the unfinished frame -- but advances the list by the committed ones only (pgsd.c:2029-2046).
    const uint64_t write_pos = h->header.index_location + sizeof(pgsd_index_entry) * h->file_index.size;
    bool ok = true;
    if (h->rank == 0)
        {
        ok = put_small(s->fd, h->frame_index.data, sizeof(pgsd_index_entry) * h->frame_index.size, write_pos);
        dev_stats().file_bytes_written += sizeof(pgsd_index_entry) * h->frame_index.size;
        }
    size_t room = h->file_index.reserved - h->file_index.size;
    size_t ncopy = h->frame_index.size < room ? h->frame_index.size : room;
    memcpy(h->file_index.data + h->file_index.size, h->frame_index.data, sizeof(pgsd_index_entry) * ncopy);
    h->file_index.size += n_commit;
    // "keep the entries of the unfinished frame": the reference copies the FIRST of them into
    // every kept slot (pgsd.c:2049-2057) -- reproduced, it decides later file contents.
    for (uint64_t i = 0; i < h->pending_index_entries; i++)
        h->frame_index.data[i] = h->frame_index.data[h->frame_index.size - h->pending_index_entries];
    h->frame_index.size = h->pending_index_entries;
    return ok ? PGSD_SUCCESS : PGSD_ERROR_IO;
    }

// Apply one recorded chunk with its K2 scan (sizes of all ranks reduced per chunk).
// ref: pgsd.c:2143-2259
int apply_op(pgsd_handle* h, PendingOp& op, const SizeScan& sc)
    {
    FileState* s = state_of(h);
    const size_t es = type_size(op.entry.type);
    if (op.n_global_auto)
        op.entry.N = (op.entry.M && es) ? sc.total / ((uint64_t)op.entry.M * es) : 0;
    const uint64_t offset_bytes = op.offset_bytes == UINT64_MAX ? sc.excl : op.offset_bytes;
    const void* host = nullptr;
    if (op.src == Src::HostUser)
        host = op.ptr;
    else if (op.src == Src::HostStage)
        host = s->stage.data() + op.stage_off;

    if (sc.maxv < h->maximum_write_buffer_size && !op.all)
        {
        // buffered: every rank appends its bytes to its own write buffer; the entry points into
        // rank 0's buffer (ref: pgsd.c:2160-2201)
        if (sc.first > h->maximum_write_buffer_size - s->root_wb_size)
            {
            int rc = flush_write_buffer(h);
            if (rc != PGSD_SUCCESS)
                return rc;
            }
        op.entry.location = (int64_t)s->root_wb_size;
        int rc = index_add(h->buffer_index, op.entry);
        if (rc != PGSD_SUCCESS)
            return rc;
        if (op.size > 0)
            {
            if (op.src == Src::Device)
                {
                std::vector<char> tmp(op.size);
                rc = dev_copy_to_host(tmp.data(), op.ptr, op.size);
                if (rc == PGSD_SUCCESS)
                    rc = bytes_append(h->write_buffer, tmp.data(), op.size);
                }
            else
                rc = bytes_append(h->write_buffer, host, op.size);
            if (rc != PGSD_SUCCESS)
                return rc;
            }
        s->root_wb_size += sc.first;
        s->wb_excl += sc.excl;
        s->wb_total += sc.total;
        }
    else
        {
        // direct: rank r's bytes go to file_size + offset_r; the file grows by the SUM of all
        // ranks' sizes even when only rank 0 writes (ref: pgsd.c:2203-2250)
        op.entry.location = (int64_t)h->file_size;
        int rc = index_add(h->frame_index, op.entry);
        if (rc != PGSD_SUCCESS)
            return rc;
        const uint64_t where = (uint64_t)h->file_size + offset_bytes;
        if ((op.all || h->rank == 0) && op.size > 0)
            {
            if (op.src == Src::Device)
                s->dev_jobs.push_back(WriteJob { op.ptr, op.size, where });
            else
                {
                if (!pwrite_all(s->fd, host, op.size, where))
                    return PGSD_ERROR_IO;
                dev_stats().file_bytes_written += op.size;
                }
            }
        h->file_size += (long long)sc.total;
        }
    h->pending_index_entries++;
    return PGSD_SUCCESS;
    }

// K2: one all-gather + scan for every chunk recorded since the last commit, then replay.
int commit_ops(pgsd_handle* h)
    {
    FileState* s = state_of(h);
    if (s->ops.empty())
        return PGSD_SUCCESS;
    const size_t n = s->ops.size();
    std::vector<uint64_t> sizes(n);
    for (size_t i = 0; i < n; i++)
        sizes[i] = s->ops[i].size;
    std::vector<SizeScan> scan(n);
    int rc = s->comm->allgather_scan(sizes.data(), scan.data(), n);
    if (rc != 0)
        return PGSD_ERROR_IO;
    for (size_t i = 0; i < n && rc == PGSD_SUCCESS; i++)
        rc = apply_op(h, s->ops[i], scan[i]);
    s->ops.clear();
    s->stage.clear();
    return rc;
    }

struct ChunkSource
    {
    Src src;
    const void* ptr; // HostUser: caller data; Device: arena pointer
    };

// shared tail of pgsd_write_chunk / pgsd_b200_write_chunk_soa: name -> id, record, maybe apply
int record_chunk(pgsd_handle* h, const char* name, int type, uint64_t N, uint32_t M, uint64_t N_global,
                 uint32_t M_global, uint64_t offset, bool all, ChunkSource cs)
    {
    FileState* s = state_of(h);
    const size_t es = type_size(type);

    // name -> id in first-seen order; new names wait in frame_names (ref: pgsd.c:2111-2141, :1340-1404)
    uint16_t id;
    std::string key(name);
    auto it = s->name_ids.find(key);
    if (it != s->name_ids.end())
        id = it->second;
    else
        {
        if (h->file_names.n_names + h->frame_names.n_names == UINT16_MAX)
            return PGSD_ERROR_NAMELIST_FULL;
        id = (uint16_t)(h->file_names.n_names + h->frame_names.n_names);
        if (!is_version2(h))
            {
            char fixed[PGSD_NAME_SIZE];
            memset(fixed, 0, sizeof(fixed));
            strncpy(fixed, name, PGSD_NAME_SIZE - 1);
            int rc = bytes_append(h->frame_names.data, fixed, PGSD_NAME_SIZE);
            if (rc != PGSD_SUCCESS)
                return rc;
            key = fixed;
            }
        else
            {
            int rc = bytes_append(h->frame_names.data, name, strlen(name) + 1);
            if (rc != PGSD_SUCCESS)
                return rc;
            }
        h->frame_names.n_names++;
        s->name_ids.emplace(key, id);
        }

    PendingOp op;
    memset(&op.entry, 0, sizeof(op.entry));
    op.entry.frame = h->cur_frame;
    op.entry.id = id;
    op.entry.type = (uint8_t)type;
    op.n_global_auto = (N_global == UINT64_MAX);
    op.entry.N = N_global;
    op.entry.M = M_global;
    op.size = N * M * es;
    op.offset_bytes = offset == PGSD_B200_OFFSET_AUTO ? UINT64_MAX : offset * es;
    op.all = all;
    op.src = cs.src;
    op.ptr = cs.ptr;
    op.stage_off = 0;

    if (s->comm->nprocs == 1)
        {
        // one rank: the scan is the identity, apply at once and write straight from the caller's memory
        SizeScan sc { 0, op.size, op.size, op.size };
        return apply_op(h, op, sc);
        }
    if (op.src == Src::HostUser)
        {
        op.src = Src::HostStage;
        op.stage_off = s->stage.size();
        if (op.size > 0)
            s->stage.insert(s->stage.end(), (const char*)cs.ptr, (const char*)cs.ptr + op.size);
        }
    s->ops.push_back(op);
    return PGSD_SUCCESS;
    }

int flush_all(pgsd_handle* h)
    {
    int rc = commit_ops(h);
    if (rc != PGSD_SUCCESS)
        return rc;
    return flush_replica(h);
    }
    } // namespace

// ------------------------------------------------------------------------------ internal API
int file_write_chunks_device(pgsd_handle* h, int n, const DeviceChunk* chunks)
    {
    if (h == nullptr || h->fh == nullptr || n < 0 || (n > 0 && chunks == nullptr))
        return PGSD_ERROR_INVALID_ARGUMENT;
    if (h->open_flags == PGSD_OPEN_READONLY)
        return PGSD_ERROR_FILE_MUST_BE_WRITABLE;
    for (int i = 0; i < n; i++)
        {
        const DeviceChunk& c = chunks[i];
        if (c.name == nullptr || c.M == 0 || (c.N > 0 && c.cols == nullptr))
            return PGSD_ERROR_INVALID_ARGUMENT;
        if (type_size(c.dst_type) == 0 || !cast_supported(c.src_type, c.dst_type))
            return PGSD_ERROR_INVALID_ARGUMENT;
        }
    FileState* s = state_of(h);
    // ONE K1 launch packs every chunk of the call into the frame arena
    std::vector<PackRequest> reqs((size_t)n);
    for (int i = 0; i < n; i++)
        {
        const DeviceChunk& c = chunks[i];
        reqs[(size_t)i] = PackRequest { c.dst_type, c.src_type, c.N, c.M, c.cols, c.host_columns, nullptr };
        }
    int rc = dev_arena_pack(reqs.data(), n, &s->dev_frame);
    if (rc != 0)
        return rc;
    s->device_frame_open = true;
    for (int i = 0; i < n; i++)
        {
        const DeviceChunk& c = chunks[i];
        rc = record_chunk(h, c.name, c.dst_type, c.N, c.M, c.N_global, c.M_global, c.offset, c.all,
                          ChunkSource { Src::Device, reqs[(size_t)i].arena_ptr });
        if (rc != PGSD_SUCCESS)
            return rc;
        }
    return PGSD_SUCCESS;
    }

int file_write_chunk_device(pgsd_handle* h, const char* name, int dst_type, uint64_t N, uint32_t M,
                            uint64_t N_global, uint32_t M_global, uint64_t offset, bool all, int src_type,
                            const Column* cols, bool host_columns)
    {
    DeviceChunk c { name, dst_type, src_type, N, M, N_global, M_global, offset, all, cols, host_columns };
    return file_write_chunks_device(h, 1, &c);
    }

int file_read_to_device(pgsd_handle* h, void* dev_dst, uint64_t bytes, uint64_t file_off)
    {
    return dev_read_file_to_device(state_of(h)->fd, dev_dst, bytes, file_off, h->open_flags == PGSD_OPEN_READONLY);
    }
} // namespace pgsdb

using namespace pgsdb;

// ================================================================================ C ABI
extern "C" {

uint32_t pgsd_make_version(unsigned int major, unsigned int minor) { return make_version(major, minor); }

bool is_root(void) { return comm()->rank == 0; }

size_t pgsd_sizeof_type(enum pgsd_type type) { return type_size((int)type); }

// ref: pgsd.c:1710-1773 + :1414-1474
int pgsd_create_and_open(struct pgsd_handle* handle, const char* fname, const char* application,
                         const char* schema, uint32_t schema_version, enum pgsd_open_flag flags,
                         int exclusive_create)
    {
    if (handle == nullptr || fname == nullptr || application == nullptr || schema == nullptr)
        return PGSD_ERROR_INVALID_ARGUMENT;
    memset(handle, 0, sizeof(*handle));
    dev_read_ahead_reset(); // ranges fetched ahead for a read-only handle may belong to the file replaced here
    if (flags == PGSD_OPEN_READONLY)
        return PGSD_ERROR_FILE_MUST_BE_WRITABLE;
    if (flags == PGSD_OPEN_READWRITE || flags == PGSD_OPEN_APPEND)
        handle->open_flags = flags;
    FileState* s = new FileState;
    s->comm = comm();
    comm_acquire();
    handle->fh = s;

    int rc = PGSD_SUCCESS;
    if (s->comm->rank == 0)
        {
        s->fd = open(fname, O_RDWR | O_CREAT | (exclusive_create ? O_EXCL : 0), 0644);
        if (s->fd < 0 || ftruncate(s->fd, 0) != 0)
            rc = PGSD_ERROR_IO;
        else
            {
            // header + 128 zero index entries + 1024 zero namelist bytes: data starts at 5376
            std::vector<char> img(sizeof(pgsd_header) + INITIAL_INDEX_SIZE * sizeof(pgsd_index_entry)
                                      + INITIAL_NAME_BUFFER_SIZE,
                                  0);
            pgsd_header hd;
            memset(&hd, 0, sizeof(hd));
            hd.magic = MAGIC_ID;
            hd.pgsd_version = make_version(2, 0);
            strncpy(hd.application, application, sizeof(hd.application) - 1);
            strncpy(hd.schema, schema, sizeof(hd.schema) - 1);
            hd.schema_version = schema_version;
            hd.index_location = sizeof(hd);
            hd.index_allocated_entries = INITIAL_INDEX_SIZE;
            hd.namelist_location = hd.index_location + sizeof(pgsd_index_entry) * hd.index_allocated_entries;
            hd.namelist_allocated_entries = INITIAL_NAME_BUFFER_SIZE / PGSD_NAME_SIZE;
            memcpy(img.data(), &hd, sizeof(hd));
            if (!pwrite_all(s->fd, img.data(), img.size(), 0))
                rc = PGSD_ERROR_IO;
            dev_stats().file_bytes_written += img.size();
            }
        }
    rc = agree(s->comm, rc);
    if (rc == PGSD_SUCCESS && s->comm->rank != 0)
        {
        s->fd = open(fname, O_RDWR);
        if (s->fd < 0)
            rc = PGSD_ERROR_IO;
        }
    if (s->comm->nprocs > 1)
        rc = agree(s->comm, rc);
    if (rc == PGSD_SUCCESS)
        rc = agree(s->comm, initialize_handle(handle));
    if (rc != PGSD_SUCCESS)
        release_state(handle);
    return rc;
    }

// ref: pgsd.c:1775-1812
int pgsd_open(struct pgsd_handle* handle, const char* fname, enum pgsd_open_flag flags)
    {
    if (handle == nullptr || fname == nullptr)
        return PGSD_ERROR_INVALID_ARGUMENT;
    memset(handle, 0, sizeof(*handle));
    if (flags != PGSD_OPEN_READWRITE && flags != PGSD_OPEN_READONLY && flags != PGSD_OPEN_APPEND)
        return PGSD_ERROR_IO; // the reference opens nothing for an unknown flag and fails on fh == NULL
    handle->open_flags = flags;
    dev_read_ahead_reset();
    FileState* s = new FileState;
    s->comm = comm();
    comm_acquire();
    handle->fh = s;
    s->fd = open(fname, flags == PGSD_OPEN_READONLY ? O_RDONLY : O_RDWR);
    int rc = s->fd < 0 ? PGSD_ERROR_IO : initialize_handle(handle);
    if (rc != PGSD_SUCCESS)
        release_state(handle);
    return rc;
    }

// ref: pgsd.c:1814-1914
int pgsd_close(struct pgsd_handle* handle)
    {
    if (handle == nullptr)
        return PGSD_ERROR_INVALID_ARGUMENT;
    if (handle->fh == nullptr)
        return PGSD_ERROR_IO;
    int rc = PGSD_SUCCESS;
    if (handle->open_flags != PGSD_OPEN_READONLY)
        {
        rc = flush_all(handle);
        int rc2 = drain_all(handle);
        if (rc == PGSD_SUCCESS)
            rc = rc2;
        if (rc != PGSD_SUCCESS)
            return rc;
        }
    dev_read_ahead_reset();
    FileState* s = state_of(handle);
    int fd = s->fd;
    s->fd = -1;
    release_state(handle);
    if (close(fd) != 0)
        return PGSD_ERROR_IO;
    return PGSD_SUCCESS;
    }

// ref: pgsd.c:1916-1953.  The frame's chunks get their file offsets here (K2).
int pgsd_end_frame(struct pgsd_handle* handle)
    {
    if (handle == nullptr || handle->fh == nullptr)
        return PGSD_ERROR_INVALID_ARGUMENT;
    if (handle->open_flags == PGSD_OPEN_READONLY)
        return PGSD_ERROR_FILE_MUST_BE_WRITABLE;
    int rc = commit_ops(handle);
    if (rc != PGSD_SUCCESS)
        return rc;
    handle->cur_frame++;
    handle->pending_index_entries = 0;
    if (handle->frame_index.size > 0 || handle->buffer_index.size > handle->index_entries_to_buffer)
        rc = flush_replica(handle);
    int rc2 = submit_device_jobs(state_of(handle));
    return rc != PGSD_SUCCESS ? rc : rc2;
    }

// ref: pgsd.c:1955-2070.  An explicit flush is also a durability point: it returns when every
// rank's queued bytes are in the file.
int pgsd_flush(struct pgsd_handle* handle)
    {
    if (handle == nullptr || handle->fh == nullptr)
        return PGSD_ERROR_INVALID_ARGUMENT;
    if (handle->open_flags == PGSD_OPEN_READONLY)
        return PGSD_ERROR_FILE_MUST_BE_WRITABLE;
    int rc = flush_all(handle);
    int rc2 = drain_all(handle);
    return rc != PGSD_SUCCESS ? rc : rc2;
    }

// ref: pgsd.c:2072-2259
int pgsd_write_chunk(struct pgsd_handle* handle, const char* name, enum pgsd_type type, uint64_t N,
                     uint32_t M, uint64_t N_global, uint32_t M_global, uint64_t offset,
                     uint64_t global_size, bool all, uint8_t flags, const void* data)
    {
    (void)global_size; // dead in the reference (pgsd.c:2147-2151)
    if (handle == nullptr || handle->fh == nullptr || name == nullptr)
        return PGSD_ERROR_INVALID_ARGUMENT;
    if (N > 0 && data == nullptr)
        return PGSD_ERROR_INVALID_ARGUMENT;
    if (M == 0)
        return PGSD_ERROR_INVALID_ARGUMENT;
    if (handle->open_flags == PGSD_OPEN_READONLY)
        return PGSD_ERROR_FILE_MUST_BE_WRITABLE;
    if (flags != 0)
        return PGSD_ERROR_INVALID_ARGUMENT;
    if (N > 0 && type_size((int)type) > 0 && dev_is_device_pointer(data))
        {
        // already packed (N, M) on the device: K1's vector-copy path moves the N*M elements (as one
        // flat column; only the byte count matters to the layout) into
        // the frame arena on the caller's stream, so `data` may be reused in stream order on return
        Column one { data, 1 };
        DeviceChunk c { name, (int)type, (int)type, N * M, 1, N_global, M_global, offset, all, &one, false };
        return file_write_chunks_device(handle, 1, &c);
        }
    return record_chunk(handle, name, (int)type, N, M, N_global, M_global, offset, all,
                        ChunkSource { Src::HostUser, data });
    }

// ref: pgsd.c:2261-2276
uint64_t pgsd_get_nframes(struct pgsd_handle* handle) { return handle ? handle->cur_frame : 0; }

// ref: pgsd.c:2279-2292
uint64_t pgsd_get_nnames(struct pgsd_handle* handle) { return handle ? handle->file_names.n_names : 0; }

// ref: pgsd.c:2295-2434
const struct pgsd_index_entry* pgsd_find_chunk(struct pgsd_handle* handle, uint64_t frame, const char* name)
    {
    if (handle == nullptr || handle->fh == nullptr || name == nullptr)
        return nullptr;
    if (frame >= handle->cur_frame)
        return nullptr;
    if (handle->open_flags != PGSD_OPEN_READONLY && flush_all(handle) != PGSD_SUCCESS)
        return nullptr;
    FileState* s = state_of(handle);
    auto it = s->name_ids.find(name);
    if (it == s->name_ids.end())
        return nullptr;
    const uint16_t id = it->second;
    const pgsd_index_entry* v = handle->file_index.data;
    if (is_version2(handle))
        {
        // whole index sorted by (frame, id): binary search
        int64_t L = 0, R = (int64_t)handle->file_index.size - 1;
        while (L <= R)
            {
            int64_t m = (L + R) / 2;
            if (v[m].frame < frame || (v[m].frame == frame && v[m].id < id))
                L = m + 1;
            else if (v[m].frame > frame || v[m].id > id)
                R = m - 1;
            else
                return v + m;
            }
        return nullptr;
        }
    // v1: frames ascend, ids inside a frame are unordered -- bisect to the last entry of the
    // frame, then scan backwards
    if (handle->file_index.size == 0)
        return nullptr;
    size_t L = 0, R = handle->file_index.size;
    do
        {
        size_t m = (L + R) / 2;
        if (frame < v[m].frame)
            R = m;
        else
            L = m;
        } while (R - L > 1);
    for (int64_t i = (int64_t)L; i >= 0 && v[i].frame == frame; i--)
        if (v[i].id == id)
            return v + i;
    return nullptr;
    }

// ref: pgsd.c:2436-2537
int pgsd_read_chunk(struct pgsd_handle* handle, void* data, const struct pgsd_index_entry* chunk,
                    uint64_t N, uint32_t M, uint32_t offset, bool all)
    {
    if (handle == nullptr || handle->fh == nullptr)
        return PGSD_ERROR_INVALID_ARGUMENT;
    if (data == nullptr || chunk == nullptr)
        return PGSD_ERROR_INVALID_ARGUMENT;
    // copy first: a flush may move the index the entry points into
    const pgsd_index_entry e = *chunk;
    if (handle->open_flags != PGSD_OPEN_READONLY)
        {
        int rc = flush_all(handle);
        int rc2 = drain_all(handle);
        if (rc != PGSD_SUCCESS || rc2 != PGSD_SUCCESS)
            return rc != PGSD_SUCCESS ? rc : rc2;
        }
    const uint64_t es = type_size(e.type);
    uint64_t size, stride = 0;
    if (!all)
        size = e.N * (uint64_t)e.M * es;
    else
        {
        // rows -> elements in 32 bits, as the reference does (pgsd.h:609, pgsd.c:2498)
        uint32_t off_elems = offset * M;
        size = N * (uint64_t)M * es;
        stride = (uint64_t)off_elems * es;
        }
    if (size == 0 || e.location == 0)
        return PGSD_ERROR_FILE_CORRUPT;
    if ((uint64_t)e.location + size + stride > (uint64_t)handle->file_size)
        return PGSD_ERROR_FILE_CORRUPT;
    FileState* s = state_of(handle);
    if (dev_is_device_pointer(data))
        return dev_read_file_to_device(s->fd, data, size, (uint64_t)e.location + stride, handle->open_flags == PGSD_OPEN_READONLY);
    int64_t got = pread_some(s->fd, data, size, (uint64_t)e.location + stride);
    if (got < 0)
        return PGSD_ERROR_IO;
    dev_stats().file_bytes_read += (uint64_t)got;
    return PGSD_SUCCESS;
    }

// ref: pgsd.c:2557-2641
const char* pgsd_find_matching_chunk_name(struct pgsd_handle* handle, const char* match, const char* prev)
    {
    if (handle == nullptr || handle->fh == nullptr || match == nullptr)
        return nullptr;
    if (handle->file_names.n_names == 0)
        return nullptr;
    if (handle->open_flags != PGSD_OPEN_READONLY && flush_all(handle) != PGSD_SUCCESS)
        return nullptr;
    const pgsd_byte_buffer& nb = handle->file_names.data;
    if (nb.data[nb.reserved - 1] != 0)
        return nullptr;
    const char* end = nb.data + nb.reserved;
    const bool v2 = is_version2(handle);
    const char* p;
    if (prev == nullptr)
        p = nb.data;
    else
        {
        if (prev < nb.data || prev >= end)
            return nullptr;
        p = v2 ? prev + strlen(prev) + 1 : prev + PGSD_NAME_SIZE;
        }
    const size_t mlen = strlen(match);
    while (p < end)
        {
        if (p[0] != 0 && strncmp(match, p, mlen) == 0)
            return p;
        p += v2 ? strlen(p) + 1 : (size_t)PGSD_NAME_SIZE;
        }
    return nullptr;
    }

// ref: pgsd.c:2643-2683
uint64_t pgsd_get_maximum_write_buffer_size(struct pgsd_handle* handle)
    {
    return handle ? handle->maximum_write_buffer_size : 0;
    }

int pgsd_set_maximum_write_buffer_size(struct pgsd_handle* handle, uint64_t size)
    {
    if (handle == nullptr || size == 0)
        return PGSD_ERROR_INVALID_ARGUMENT;
    // recorded chunks were written under the old limit
    if (handle->fh != nullptr && handle->open_flags != PGSD_OPEN_READONLY)
        {
        int rc = commit_ops(handle);
        if (rc != PGSD_SUCCESS)
            return rc;
        }
    handle->maximum_write_buffer_size = size;
    return PGSD_SUCCESS;
    }

uint64_t pgsd_get_index_entries_to_buffer(struct pgsd_handle* handle)
    {
    return handle ? handle->index_entries_to_buffer : 0;
    }

int pgsd_set_index_entries_to_buffer(struct pgsd_handle* handle, uint64_t number)
    {
    if (handle == nullptr || number == 0)
        return PGSD_ERROR_INVALID_ARGUMENT;
    handle->index_entries_to_buffer = number;
    return PGSD_SUCCESS;
    }

// ref: pgsd.c:152-172
void pgsd_bcast_index_entry(struct pgsd_index_entry* e)
    {
    Comm* c = comm();
    if (e == nullptr || c->nprocs == 1)
        return;
    uint64_t w[4];
    memcpy(w, e, 32);
    std::vector<uint64_t> all((size_t)c->nprocs * 4);
    if (c->allgather(w, all.data(), 4) == 0)
        memcpy(e, all.data(), 32);
    }

} // extern "C"
