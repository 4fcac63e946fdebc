python -m pytest tests -m gpu -x -q -k "hilbert" > gpurun_out/pytest_gpu.log 2>&1; tail -5 gpurun_out/pytest_gpu.log
python bench_configs.py --config c4 --no-cpu 2>/dev/null | cut -c1-260
FFTCONV_CUDA_NO_PERSISTENT_HILBERT=1 python bench_configs.py --config c4 --no-cpu 2>/dev/null | cut -c1-200
