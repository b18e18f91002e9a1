set -x
mkdir -p gpurun_out
: > gpurun_out/ccl_ab.jsonl
for v in default ccl_old; do
  if [ $v = default ]; then L=$PWD/kimimaro_b200/libb2t.so; else L=$PWD/kimimaro_b200/_variants/$v.so; fi
  B2T_LIB=$L B2T_X=$v timeout 300 python scripts/ccl_time.py 512 5 >> gpurun_out/ccl_ab.jsonl 2>> gpurun_out/ccl_ab.err
done
cat gpurun_out/ccl_ab.jsonl; tail -3 gpurun_out/ccl_ab.err
