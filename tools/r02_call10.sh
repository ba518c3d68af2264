set -x
timeout 300 python -m pytest tests/test_gpu_gp.py -m gpu -q -x > gpurun_out/r02_c10_tests.log 2>&1; echo rc=$? >> gpurun_out/r02_c10_tests.log
( time timeout 900 python bench.py --steps 20 --warmup 3 --no-predict --no-small --no-s4 --no-cpu-baseline ) > gpurun_out/r02_c10_bench.log 2> gpurun_out/r02_c10_bench.err
tail -n 3 gpurun_out/r02_c10_tests.log; tail -n 1 gpurun_out/r02_c10_bench.log | cut -c1-200; tail -n 4 gpurun_out/r02_c10_bench.err
