set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/gpu_tests.log 2>&1; tail -3 gpurun_out/gpu_tests.log
timeout 600 python bench.py --gpus 1 --steps 8 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -c 300 gpurun_out/bench_n1.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r02_bench_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-strict > gpurun_out/ncu_bench.log 2>&1
tail -c 200 gpurun_out/ncu_bench.log
timeout 600 ncu --set full --clock-control none --import-source on -k 'regex:trace_kernel|edf_label_kernel|edf_multi_kernel' -c 4 -f -o /tmp/r02_trace python scripts/phase_times.py 512 1 > gpurun_out/ncu_full_a.log 2>&1
tail -c 200 gpurun_out/ncu_full_a.log
ncu -i /tmp/r02_trace.ncu-rep --page raw --csv > gpurun_out/r02_trace_edf_raw.csv 2> /dev/null
ls -la /tmp/r02_trace.ncu-rep
timeout 600 ncu --set full --clock-control none --import-source on -k 'regex:edt_pass|ccl_merge_kernel|ccl_init_kernel|ccl_flatten' -c 8 -f -o /tmp/r02_k1 python scripts/phase_times.py 512 1 > gpurun_out/ncu_full_b.log 2>&1
tail -c 200 gpurun_out/ncu_full_b.log
ncu -i /tmp/r02_k1.ncu-rep --page raw --csv > gpurun_out/r02_k1_ccl_raw.csv 2> /dev/null
ls -la /tmp/r02_k1.ncu-rep
gzip -c /tmp/r02_k1.ncu-rep > /tmp/k1.gz; gzip -c /tmp/r02_trace.ncu-rep > /tmp/tr.gz; ls -la /tmp/*.gz
[ $(stat -c %s /tmp/k1.gz) -lt 12000000 ] && cp /tmp/k1.gz gpurun_out/r02_k1_ccl.ncu-rep.gz
[ $(stat -c %s /tmp/tr.gz) -lt 12000000 ] && cp /tmp/tr.gz gpurun_out/r02_trace_edf.ncu-rep.gz
timeout 900 python bench.py --size 1024 --steps 2 --warmup 1 --no-cpu --no-strict > gpurun_out/bench_1024.json 2> gpurun_out/bench_1024.err; tail -c 300 gpurun_out/bench_1024.err; cut -c1-300 gpurun_out/bench_1024.json
du -sh gpurun_out
