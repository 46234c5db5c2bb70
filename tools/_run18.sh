mkdir -p gpurun_out/ev18
python -m pytest tests -x -q -m gpu -k "quality_filter_with_sparse" > gpurun_out/ev18/pytest_q.log 2>&1; tail -5 gpurun_out/ev18/pytest_q.log
python -m pytest tests -x -q -m gpu > gpurun_out/ev18/pytest_all.log 2>&1; tail -3 gpurun_out/ev18/pytest_all.log
python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu > gpurun_out/ev18/bench_main.json 2> gpurun_out/ev18/bench_main.err
python bench.py --steps 5 --warmup 3 --no-e2e --qual 13.5 > gpurun_out/ev18/bench_q_sparse.json 2> gpurun_out/ev18/bench_q_sparse.err
python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu --qual 13.5 --scan-dense --reads 20000 > gpurun_out/ev18/bench_q_dense.json 2> gpurun_out/ev18/bench_q_dense.err
python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu --qual 13.5 --reads 20000 > gpurun_out/ev18/bench_q_sparse20k.json 2> gpurun_out/ev18/bench_q_sparse20k.err
tail -c 400 gpurun_out/ev18/*.err
for f in gpurun_out/ev18/bench_*.json; do echo $f; python -c "
import json,sys
d=json.loads(open('$f').read().strip().split('\n')[-1])
print(d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'], d.get('parity_check') and d['parity_check'].get('equal'), d['config']['workload'])
"; done
