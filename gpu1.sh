mkdir -p gpurun_out
timeout 600 python bench.py --workload radon100k --eps-mode per-chain --chains 592 --burn-in 20 --draws 10 --step-size 0.001 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/t14_big.json 2> gpurun_out/t14_big.err; echo rc=$?
python -c "
import json
d=json.loads(open('gpurun_out/t14_big.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['steps_detail'])"
timeout 600 python -m pytest tests -m gpu -q -x -k "large" 2>&1 | tail -3
