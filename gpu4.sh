mkdir -p gpurun_out
( time python bench.py > gpurun_out/r1c_bench_n1.json 2> gpurun_out/r1c_bench_n1.err ) 2>&1 | tail -3; echo "bench rc=$?"
python -c "
import json
d=json.loads(open('gpurun_out/r1c_bench_n1.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e'], d['gpu_launches'], d['diagnostics'], d['roofline']['frac'], d['cpu_baseline'], d['ess_per_sec'])"
( time python bench.py --impl reference > gpurun_out/r1c_ref_n1.json 2> gpurun_out/r1c_ref_n1.err ) 2>&1 | tail -3
tail -c 400 gpurun_out/r1c_ref_n1.json
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -3
