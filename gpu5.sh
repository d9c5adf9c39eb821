mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_multi_gpu.py -q -x 2>&1 | tail -15 | cut -c1-300
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 2 --warmup 2 --no-cpu-baseline > gpurun_out/r1d_bench_n2.json 2> gpurun_out/r1d_bench_n2.err; echo rc=$?
tail -5 gpurun_out/r1d_bench_n2.err
python -c "
import json
d=json.loads(open('gpurun_out/r1d_bench_n2.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['n_gpus'], d['e2e']['value'], d['config']['parallelism'])"
