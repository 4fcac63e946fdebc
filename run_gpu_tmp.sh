python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py > gpurun_out/bench_v12.log 2>gpurun_out/bench_v12.err; python -c "
import json; d=json.loads(open('gpurun_out/bench_v12.log').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['value'], d['roofline']['frac'], d['e2e']['ms_per_step'], d['e2e']['value'], d['gpu_launches'], d['cpu_baseline']['value'])"
python bench.py --impl reference > gpurun_out/bench_v12_ref.log 2>gpurun_out/bench_v12_ref.err; tail -c 300 gpurun_out/bench_v12_ref.log
python bench_configs.py --config all > gpurun_out/configs_v12.log 2>gpurun_out/configs_v12.err
python bench_configs.py --config rf --no-cpu >> gpurun_out/configs_v12.log 2>>gpurun_out/configs_v12.err
cut -c1-170 gpurun_out/configs_v12.log
