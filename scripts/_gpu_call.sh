set -x
mkdir -p gpurun_out
B2T_LIB=kimimaro_b200/_variants/prof.so timeout 300 python scripts/trace_prof.py > gpurun_out/trace_prof.jsonl 2> gpurun_out/trace_prof.err
tail -3 gpurun_out/trace_prof.err
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_b.json 2> gpurun_out/bench_b.err
tail -3 gpurun_out/bench_b.err
