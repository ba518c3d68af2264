set -x
timeout 600 python -m pytest tests -m gpu -q -s > gpurun_out/r02_c6_tests.log 2>&1; echo rc=$? >> gpurun_out/r02_c6_tests.log
timeout 300 python tools/predict_routes.py > gpurun_out/r02_c6_routes.log 2>&1
for v in base exp10; do
  GPRB_LIB=variants/libgpr_b200_$v.so timeout 200 python tools/perf_s5.py 340 1 2 >> gpurun_out/r02_c6_perf.log 2>&1
done
( time timeout 900 python bench.py --steps 2 --warmup 1 --predict-structures 640 --s4-budget-s 40 ) > gpurun_out/r02_c6_bench.log 2> gpurun_out/r02_c6_bench.err
grep -E "passed|failed|FAILED|sigma routes|Error" gpurun_out/r02_c6_tests.log | tail -n 40; cat gpurun_out/r02_c6_routes.log gpurun_out/r02_c6_perf.log | cut -c1-330; tail -n 4 gpurun_out/r02_c6_bench.err
