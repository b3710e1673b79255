// file_stage.h -- host side of K3: how one staged piece of chunk bytes reaches the file.  No CUDA here, so the
// CPU test-suite can hammer it (tests/test_file_stage.py) and tools can time it without a GPU.
#pragma once
#include <cstdint>

namespace pgsdb
{
enum class FileMode : int
    {
    Auto = 0,   // mapping path on tmpfs, pwrite everywhere else (PGSD_B200_FILE_MODE unset)
    Pwrite = 1, // always pwrite
    Mmap = 2    // mapping path on every file system that supports it (development)
    };
FileMode file_mode_from_env();
uint64_t file_page_size();
bool file_is_tmpfs(int fd);

// Writes p[0, len) at file offset off.  use_mmap: the WHOLE pages inside [off, off + len) are copied through a
// short-lived shared mapping of their own (after the range was allocated with fallocate, so a full file system is
// reported as an error here and not as SIGBUS at the store); the partial pages at either end go through pwrite.
// No page is ever touched by two mappings or by a mapping and a pwrite: see DESIGN.md section 4.
// Returns false on an I/O error with *left = bytes not written.
bool file_write_piece(int fd, const char* p, uint64_t off, uint64_t len, bool use_mmap, uint64_t* left);

// Length of the first piece of a job that starts at file offset `off`, so that every following piece of `piece`
// bytes (a multiple of the page size) starts on a page boundary.
uint64_t file_first_piece_len(uint64_t off, uint64_t bytes, uint64_t piece);

// Host-only ceiling of the file stage on the file's target: `threads` threads put `bytes` bytes (from one buffer
// of `piece` bytes each) at [off, off + bytes) of fd with file_write_piece.  Returns seconds, < 0 on error.
double file_stage_ceiling(int fd, uint64_t off, uint64_t bytes, uint64_t piece, int threads, bool use_mmap);
} // namespace pgsdb
