mkdir -p gpurun_out/ev19
python -m pytest tests -x -q -m gpu > gpurun_out/ev19/pytest_all.log 2>&1; tail -3 gpurun_out/ev19/pytest_all.log
for i in 1 2; do python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read());print('base', d['ms_per_step'],d['roofline']['kernel_ms'],d['roofline']['frac'])"; done
cp mcaller_b200/libmcaller_b200.so /tmp/lib_orig.so
cp build/variants/lib_w_swz.so mcaller_b200/libmcaller_b200.so
for i in 1 2; do python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read());print('swz', d['ms_per_step'],d['roofline']['kernel_ms'],d['roofline']['frac'])"; done
python -m pytest tests -x -q -m gpu -k "scan_runs or layout_fuzz or odd_line" 2>&1 | tail -2
cp /tmp/lib_orig.so mcaller_b200/libmcaller_b200.so
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/ev19/launches.csv python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu > /dev/null 2>&1
python tools/launch_summary.py gpurun_out/ev19/launches.csv | head -8
