mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/t8_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t8_pytest.log
tail -4 gpurun_out/t8_pytest.log
for cfg in "2 512" "4 768"; do
  set -- $cfg
  echo "=== T=$1 MAXT=$2"
  PM4_TEAM_WARPS=$1 PM4_MAXT=$2 timeout 300 python bench.py --steps 2 --warmup 2 --no-cpu-baseline --no-e2e 2> gpurun_out/t8_bench_$1_$2.err | tee gpurun_out/t8_bench_$1_$2.json | python -c "
import sys, json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l[:300]); continue
    print(d['value'], d['ms_per_step'], d['diagnostics']['mean_leapfrogs_per_draw'], d['diagnostics']['max_split_rhat'])
"
  tail -3 gpurun_out/t8_bench_$1_$2.err
done
