#!/bin/bash
# tuning helper: times bench.py with alternative builds of the library (MC_SCAN_WARPS / MC_SCAN_MIN_CTAS)
cd /root/repo
cp mcaller_b200/libmcaller_b200.so /tmp/lib_orig.so
for f in build/variants/lib_w*.so; do
  cp $f mcaller_b200/libmcaller_b200.so
  touch mcaller_b200/libmcaller_b200.so
  echo "== $f"
  timeout 300 python bench.py --reads 30000 --steps 3 --warmup 3 --no-e2e --no-cpu 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read());print(d['ms_per_step'],d['roofline']['kernel_ms'],d['roofline']['frac'])"
done
cp /tmp/lib_orig.so mcaller_b200/libmcaller_b200.so
