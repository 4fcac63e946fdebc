python bench_configs.py --config c1 --no-cpu 2>&1 | tail -1 | cut -c1-330
python -m pytest tests -m gpu -x -q -k "multi or device or cuda_array or dlpack or small_jobs or golden" 2>&1 | tail -2
