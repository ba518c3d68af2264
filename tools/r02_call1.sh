set -x
python -m pytest tests -m gpu -x -q > gpurun_out/r02_t_default.log 2>&1; echo rc=$? >> gpurun_out/r02_t_default.log
GPRB_POTRF_LOWER=1 python -m pytest tests -m gpu -x -q > gpurun_out/r02_t_lower.log 2>&1; echo rc=$? >> gpurun_out/r02_t_lower.log
GPRB_TEST_TWO_STAGE=1 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k two_stage > gpurun_out/r02_t_two.log 2>&1; echo rc=$? >> gpurun_out/r02_t_two.log
GPRB_KFF_TWO_STAGE=1 timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/r02_t_two_all.log 2>&1; echo rc=$? >> gpurun_out/r02_t_two_all.log
python tools/perf_one.py 4000 32 32 0 0 3 > gpurun_out/r02_perf_block.log 2>&1
GPRB_KFF_TWO_STAGE=1 timeout 120 python tools/perf_one.py 4000 32 32 0 0 3 > gpurun_out/r02_perf_two.log 2>&1
python tools/perf_one.py 4000 32 32 0 3 3 >> gpurun_out/r02_perf_block.log 2>&1
GPRB_KFF_TWO_STAGE=1 timeout 120 python tools/perf_one.py 4000 32 32 0 3 3 >> gpurun_out/r02_perf_two.log 2>&1
python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-predict > gpurun_out/r02_bench_upper.log 2>&1
GPRB_POTRF_LOWER=1 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-predict > gpurun_out/r02_bench_lower.log 2>&1
tail -3 gpurun_out/r02_t_*.log gpurun_out/r02_perf_*.log
