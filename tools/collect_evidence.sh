#!/bin/bash
# Re-collects the round evidence under gpurun_out/ev/ on a GPU box (run via gpurun from the repo root):
#   1. default bench.py line (the number the driver re-measures)        -> bench_default.json
#   2. ncu launch list of one full-size step                             -> launches_full_step.csv
#   3. ncu DRAM bytes of the full-size k_scan launch                     -> k_scan_dram_full_size.csv
#   4. ncu --set full of k_scan on a 2 GB input (+ source page)          -> k_scan_full.ncu-rep
#   5. ncu --set full of the secondary kernels at full size              -> secondary_full.ncu-rep
#   6. compute-sanitizer memcheck over parity tests                      -> memcheck.txt
# Summaries are produced afterwards in the build container with tools/ncu_summary.py, tools/ncu_phases.py and
# tools/launch_summary.py and copied into profiles/.
set -u
O=gpurun_out/ev
mkdir -p $O
timeout 600 python bench.py > $O/bench_default.json 2> $O/bench_default.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_full_step.csv \
    python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu > $O/launches.log 2>&1
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
    -k k_scan --launch-skip 3 -c 1 --csv --log-file $O/k_scan_dram_full_size.csv \
    python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu > $O/dram.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k k_scan --launch-skip 3 -c 1 -f -o $O/k_scan_full \
    python bench.py --reads 4000 --steps 1 --warmup 3 --no-e2e --no-cpu > $O/full.log 2>&1
# the launches of the step after the three warm-up steps (6 matching launches per step)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_gather_finish|k_mlp|k_windows|k_seg_fix|k_first_m' \
    --launch-skip 18 -c 6 -f -o $O/secondary_full python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu > $O/secondary.log 2>&1
# compute-sanitizer (memcheck, racecheck, synccheck) over the golden, quiet-chunk / distant-closer, odd-shape, carry and -q read-first tests
SEL="diffs_match and (gatc_s1 or adversarial or A_s2) or quiet_chunks and junk or odd_number_shapes or chunked_equals_whole and gatc_s1 or worker_ranges and gatc_s1 or quality_filter and (plain or names) and GAT-0"
for tool in memcheck racecheck synccheck; do
  ( timeout 1500 compute-sanitizer --tool $tool --error-exitcode 9 python -m pytest tests/test_gpu_golden.py tests/test_gpu_kernels.py -x -q -k "$SEL" 2>&1 | tail -4
    echo "exit=$?  command: compute-sanitizer --tool $tool --error-exitcode 9 python -m pytest tests/test_gpu_golden.py tests/test_gpu_kernels.py -x -q -k '$SEL'" ) > $O/sanitizer_$tool.txt
done
ls -la $O
tail -c 600 $O/bench_default.json
cat $O/sanitizer_*.txt
