python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -5 gpurun_out/pytest_gpu.log
python bench_configs.py --config all --no-cpu > gpurun_out/configs_r2.log 2>/dev/null; cut -c1-330 gpurun_out/configs_r2.log
