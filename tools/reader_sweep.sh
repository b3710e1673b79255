for t in 8 12; do echo "threads=$t"; PGSD_B200_READER_THREADS=$t timeout 200 python tools/profile_read.py 2>&1 | tail -6; done
