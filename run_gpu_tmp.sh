python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -12 gpurun_out/pytest_gpu.log
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_v9.log 2>gpurun_out/bench_v9.err; python -c "
import json; d=json.loads(open('gpurun_out/bench_v9.log').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['value'], d['e2e'], d['roofline']['frac'])"
