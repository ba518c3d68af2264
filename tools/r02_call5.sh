set -x
python -m pytest tests -m gpu -q -s > gpurun_out/r02_c5_tests.log 2>&1; echo rc=$? >> gpurun_out/r02_c5_tests.log
python tools/predict_routes.py > gpurun_out/r02_c5_routes.log 2>&1
for ph in so3 pack kef kee lml predict; do
  ncu --set full --import-source on --clock-control none -o gpurun_out/r02_misc_$ph -f python tools/profile_misc.py $ph > gpurun_out/r02_c5_ncu_$ph.log 2>&1
done
( time python bench.py --steps 3 --warmup 3 --s4-budget-s 60 ) > gpurun_out/r02_c5_bench.log 2> gpurun_out/r02_c5_bench.err
grep -E "passed|failed|FAILED|sigma routes" gpurun_out/r02_c5_tests.log | tail -n 30; cat gpurun_out/r02_c5_routes.log; tail -n 4 gpurun_out/r02_c5_bench.err
