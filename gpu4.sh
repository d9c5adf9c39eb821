mkdir -p gpurun_out
python bench.py > gpurun_out/r1b_bench_n1.json 2> gpurun_out/r1b_bench_n1.err; echo "bench rc=$?"
tail -c 1500 gpurun_out/r1b_bench_n1.json
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r1b_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r1b_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r1b_ncu_l.log 2>&1
python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/r1b_plain2.log 2>&1 && ncu --set full --import-source on --clock-control none -k regex:nuts_kernel -s 4 -c 1 -o gpurun_out/r1b_full -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/r1b_ncu_f.log 2>&1
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')"
