python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -5 gpurun_out/pytest_gpu.log
python bench_configs.py --config rf --no-cpu > gpurun_out/rf.log 2>gpurun_out/rf.err; cat gpurun_out/rf.log; tail -3 gpurun_out/rf.err
python bench_configs.py --config c4 --no-cpu 2>/dev/null | cut -c1-400
python bench.py > gpurun_out/bench_e2e.log 2>gpurun_out/bench_e2e.err; cut -c1-1500 gpurun_out/bench_e2e.log
