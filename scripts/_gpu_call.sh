mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/gpu_tests.log 2>&1; tail -2 gpurun_out/gpu_tests.log; grep -E "FAILED|Error" gpurun_out/gpu_tests.log | head -5
