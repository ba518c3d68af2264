set -x
python -m pytest tests -m gpu -x -q > gpurun_out/r02_c2_tests.log 2>&1; echo rc=$? >> gpurun_out/r02_c2_tests.log
for v in base staged; do
  GPRB_LIB=variants/libgpr_b200_$v.so python tools/perf_s5.py 340 1 2 >> gpurun_out/r02_c2_perf.log 2>&1
  GPRB_LIB=variants/libgpr_b200_$v.so python tools/perf_s5.py 100 1 3 >> gpurun_out/r02_c2_perf.log 2>&1
done
python tools/perf_s5.py 340 0 2 >> gpurun_out/r02_c2_perf.log 2>&1
GPRB_KFF_TWO_STAGE=0 python tools/perf_s5.py 340 0 2 >> gpurun_out/r02_c2_perf.log 2>&1
python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r02_c2_bench.log 2>&1
cat gpurun_out/r02_c2_perf.log; tail -n 3 gpurun_out/r02_c2_tests.log
