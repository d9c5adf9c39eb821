mkdir -p gpurun_out
for cfg in "2 512" "4 768"; do
  set -- $cfg
  PM4_TEAM_WARPS=$1 PM4_MAXT=$2 timeout 600 ncu --metrics gpu__time_duration.sum,sm__cycles_active.avg,sm__cycles_elapsed.max,smsp__inst_executed.sum,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:nuts_kernel --csv --log-file gpurun_out/t3_ncu_$1_$2.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/t3_ncu_$1_$2.log 2>&1
done
