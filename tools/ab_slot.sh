# A/B timings of the scatter's launch shape: shared-memory carve-out (KB; 0 = the driver's choice) x tile x rows per
# thread.  One process per carve-out (that switch is read once).  Usage: bash tools/ab_slot.sh > out.txt
for carve in 0 100 132 164 196 228; do
  echo "== PGSD_B200_SLOT_CARVEOUT=$carve"
  PGSD_B200_SLOT_CARVEOUT=$carve TIME_SLOT_ONLY="slot path,slot per" python tools/time_slot.py 16777216 2>&1 | grep "perm=0" | sed 's/census 0.00[0-9] //; s/pairs 0.00[0-9] //' | cut -c1-170
done
