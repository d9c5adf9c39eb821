mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 2 --warmup 2 --no-cpu-baseline > gpurun_out/r1d_bench_n8.json 2> gpurun_out/r1d_bench_n8.err; echo rc=$?
tail -5 gpurun_out/r1d_bench_n8.err
python -c "
import json
d=json.loads(open('gpurun_out/r1d_bench_n8.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['n_gpus'], d['e2e']['value'], d['diagnostics'])"
