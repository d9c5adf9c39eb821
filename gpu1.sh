mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -q -x -k "glm or tcgen05 or dense" 2>&1 | tail -3
timeout 300 ncu --metrics gpu__time_duration.sum,sm__inst_executed.avg.per_cycle_active --clock-control none -k regex:glm --csv --log-file gpurun_out/t15_glm.csv python gpurun_glm3.py > gpurun_out/t15_glm.log 2>&1
tail -2 gpurun_out/t15_glm.log
