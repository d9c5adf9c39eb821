mkdir -p gpurun_out
python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/r1c_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r1c_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/r1c_ncu_l.log 2>&1
echo "launch list rc=$?"
python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-e2e --burn-in 30 --draws 30 > gpurun_out/r1c_plain2.log 2>&1 && ncu --set full --import-source on --clock-control none -k regex:nuts_kernel -s 31 -c 1 -o gpurun_out/r1c_full_draws -f python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-e2e --burn-in 30 --draws 30 > gpurun_out/r1c_ncu_f.log 2>&1
echo "full rc=$?"
ls -la gpurun_out/r1c*
