mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x -k "large" > gpurun_out/t9_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t9_pytest.log
tail -30 gpurun_out/t9_pytest.log | cut -c1-400
