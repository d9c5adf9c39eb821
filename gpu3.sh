mkdir -p gpurun_out
PM4_TEAM_WARPS=2 PM4_MAXT=512 timeout 900 ncu --set full --import-source on --clock-control none -k regex:nuts_kernel -c 1 -o gpurun_out/t6_full_2_512 -f python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-e2e --burn-in 30 --draws 10 > gpurun_out/t6_full.log 2>&1
ls -la gpurun_out/*.ncu-rep
