mkdir -p gpurun_out
timeout 1500 python scripts/gpu_parity_campaign.py 40 3 > gpurun_out/campaign3.jsonl 2> gpurun_out/campaign3.err; echo rc=$?; tail -3 gpurun_out/campaign3.err; tail -1 gpurun_out/campaign3.jsonl | cut -c1-400; grep '"dense": true' gpurun_out/campaign3.jsonl | cut -c1-330 | head -8
