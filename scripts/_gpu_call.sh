set -x
mkdir -p gpurun_out
timeout 1500 python scripts/gpu_parity_campaign.py 160 2 > gpurun_out/campaign2.jsonl 2> gpurun_out/campaign2.err; echo rc=$?; tail -3 gpurun_out/campaign2.err; tail -1 gpurun_out/campaign2.jsonl | cut -c1-400; grep '": false' gpurun_out/campaign2.jsonl | cut -c1-500 | head -5
timeout 600 python -m pytest tests/test_campaign_gpu.py -x -q -m gpu 2>&1 | tail -2
