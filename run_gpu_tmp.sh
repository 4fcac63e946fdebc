python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -8 gpurun_out/pytest_gpu.log
python bench_configs.py --config ksweep --no-cpu > gpurun_out/ksweep.log 2>gpurun_out/ksweep.err; cat gpurun_out/ksweep.log | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(d.get('dtype'),d.get('k'),d.get('fft_size'),round(d.get('ms_per_step',0),3),round(d.get('frac_of_measured_hbm_peak',0),3),d.get('error',''))
"
python bench_configs.py --config c3 --no-cpu > gpurun_out/c3.log 2>/dev/null; cut -c1-300 gpurun_out/c3.log
