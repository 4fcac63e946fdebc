python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
python bench.py > gpurun_out/bench_v11.log 2>gpurun_out/bench_v11.err; tail -c 600 gpurun_out/bench_v11.log
python bench.py --impl reference > gpurun_out/bench_v11_ref.log 2>gpurun_out/bench_v11_ref.err; tail -c 500 gpurun_out/bench_v11_ref.log
python bench_configs.py --config all > gpurun_out/configs_v11.log 2>gpurun_out/configs_v11.err; cut -c1-160 gpurun_out/configs_v11.log
python bench_configs.py --config rf --no-cpu >> gpurun_out/configs_v11.log 2>>gpurun_out/configs_v11.err
python bench_configs.py --config ksweep --no-cpu > gpurun_out/ksweep_v11.log 2>gpurun_out/ksweep_v11.err; wc -l gpurun_out/ksweep_v11.log
CMD="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e"
$CMD > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_v11.csv $CMD > gpurun_out/ncu_l11.log 2>&1
$CMD > gpurun_out/plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:oa_n2048 -s 1 -c 1 -o /tmp/prof_v11 -f $CMD > gpurun_out/ncu_v11.log 2>&1
ncu -i /tmp/prof_v11.ncu-rep --page details > gpurun_out/v11_details.txt 2>/dev/null
ncu -i /tmp/prof_v11.ncu-rep --page raw --csv > gpurun_out/v11_raw.csv 2>/dev/null
ncu -i /tmp/prof_v11.ncu-rep --page source --csv > gpurun_out/v11_source.csv 2>/dev/null
du -sh gpurun_out
