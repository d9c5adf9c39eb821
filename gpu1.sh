mkdir -p gpurun_out
for cfg in "100 100" "1000 1000"; do set -- $cfg
timeout 900 python bench.py --eps-mode pooled --burn-in $1 --draws $2 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/t13_pooled_$1.json 2> gpurun_out/t13_pooled_$1.err; echo rc=$?
python -c "
import json
d=json.loads(open('gpurun_out/t13_pooled_$1.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches'], d['diagnostics'], d['steps_detail'], d['ess_per_sec'])"
tail -3 gpurun_out/t13_pooled_$1.err
done
