for sr in 2 3 4 6 10; do for rows in 4096 6144 8192; do
  BCU_BIN_SUBROWS=$sr BCU_BIN_ROWS=$rows timeout 300 python bench.py --no-also --no-e2e --no-cpu-baseline --steps 5 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('subrows $sr rows $rows  ms %.2f  tiles %d  bytes %.0f MB build %.1f ms'%(d['ms_per_step'], d['config']['index']['binned_tiles'], d['config']['index']['device_bytes']/1e6, d['build']['ms']))"
done; done
