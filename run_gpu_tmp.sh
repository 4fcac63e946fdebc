CMD="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e"
$CMD > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_v10.csv $CMD > gpurun_out/ncu_l10.log 2>&1
tail -1 gpurun_out/ncu_l10.log | cut -c1-100
for C in c3:oa_n2048 c4:hilbert_pk; do
  cfg=${C%%:*}; pat=${C##*:}
  CMDC="python bench_configs.py --config $cfg --no-cpu --steps 2 --warmup 1"
  $CMDC > gpurun_out/plain_$cfg.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:$pat -s 1 -c 1 -o /tmp/prof_$cfg -f $CMDC > gpurun_out/ncu_${cfg}_v10.log 2>&1
  ncu -i /tmp/prof_$cfg.ncu-rep --page details > gpurun_out/${cfg}_v10_details.txt 2>/dev/null
  ncu -i /tmp/prof_$cfg.ncu-rep --page raw --csv > gpurun_out/${cfg}_v10_raw.csv 2>/dev/null
  ncu -i /tmp/prof_$cfg.ncu-rep --page source --csv > gpurun_out/${cfg}_v10_source.csv 2>/dev/null
  ls -la /tmp/prof_$cfg.ncu-rep
done
du -sh gpurun_out
