python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -5 gpurun_out/pytest_gpu.log
python bench_configs.py --config all --no-cpu > gpurun_out/configs_pk.log 2> gpurun_out/configs_pk.err; cut -c1-260 gpurun_out/configs_pk.log
