python bench.py > gpurun_out/bench_r1_final.log 2> gpurun_out/bench_r1_final.err; cut -c1-200 gpurun_out/bench_r1_final.log
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r1_ref.log 2> gpurun_out/bench_r1_ref.err; cut -c1-300 gpurun_out/bench_r1_ref.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_v8.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_l.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:oa_n2048 -c 1 -s 2 -o gpurun_out/prof_v8 -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_v8.log 2>&1
tail -1 gpurun_out/ncu_v8.log | cut -c1-80
