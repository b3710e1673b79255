#!/bin/bash
# host-side ceiling of the file stage on this box (page cache / tmpfs); see pwrite_bench.cpp
D=${1:-/dev/shm}
B=$(dirname "$0")/pwrite_bench
for mode in pwrite mmap; do
  for t in 1 2 4 8 12 16; do $B $D/pwb.bin 8192 16 $t $mode; done
done
$B $D/pwb.bin 8192 4 16 pwrite
$B $D/pwb.bin 8192 64 16 pwrite
$B $D/pwb.bin 8192 16 16 falloc
