#!/bin/bash
# per-piece variants of the file stage (what a writer thread can do with one 16 MiB piece); see pwrite_bench.cpp
D=${1:-/dev/shm}
B=$(dirname "$0")/pwrite_bench
for mode in mmap_piece mmap_piece_pop falloc_piece falloc_piece_pop pwrite; do
  for t in 4 8 16 32; do $B $D/pwb.bin 8192 16 $t $mode; done
done
for mode in mmap_piece falloc_piece_pop; do
  for p in 4 64; do $B $D/pwb.bin 8192 $p 8 $mode; done
done
