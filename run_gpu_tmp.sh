python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -4 gpurun_out/pytest_gpu.log
python bench_configs.py --config ksweep --no-cpu > gpurun_out/ksweep_v12.log 2>gpurun_out/ksweep_v12.err; cat gpurun_out/ksweep_v12.log | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(d.get('dtype'),d.get('k'),d.get('fft_size'),round(d.get('ms_per_step',0),3),round(d.get('frac_of_measured_hbm_peak',0),3),d.get('error',''))
"
python bench_configs.py --config c1 --no-cpu 2>/dev/null | cut -c1-200
