cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests -x -q -m gpu -k "dense or tcgen05" 2>&1 | tail -5
timeout 600 python bench.py --workload logit --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/logit2.json 2>gpurun_out/logit2.err; echo rc=$?
cut -c1-200 gpurun_out/logit2.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/logit_launches2.csv python bench.py --workload logit --steps 1 --warmup 0 --burn-in 2 --draws 2 --no-cpu-baseline --no-e2e > gpurun_out/logit_ncu2.log 2>&1
echo rc=$?
