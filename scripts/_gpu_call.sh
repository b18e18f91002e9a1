mkdir -p gpurun_out
timeout 300 python scripts/k1_shot.py 7 > gpurun_out/k1_shot.log 2>&1; tail -12 gpurun_out/k1_shot.log | cut -c1-300
