set -x
python -m pytest tests -m gpu -x -q -s > gpurun_out/r02_c4_tests.log 2>&1; echo rc=$? >> gpurun_out/r02_c4_tests.log
( time python bench.py ) > gpurun_out/r02_c4_bench.log 2> gpurun_out/r02_c4_bench.err
tail -n 30 gpurun_out/r02_c4_tests.log; tail -n 5 gpurun_out/r02_c4_bench.err
