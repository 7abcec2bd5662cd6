#!/usr/bin/env bash
# A/B of gather kernel builds (profiles/_variants/*.so) through the bench harness: kernel ms per 16-batch chunk and patches/s
for v in "$@"; do
  for rep in 1 2; do
    DEEPHISTO_B200_LIB=$PWD/profiles/_variants/$v.so python bench.py --steps 320 --warmup 32 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$v', round(d['roofline']['kernel_ms_avg_full_chunk']*1e3,1), 'us/chunk', round(d['value']/1e6,3), 'Mpatches/s', round(d['roofline']['frac'],3))"
  done
done
