set -x
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_final_smoke.log 2>&1; echo rc=$? >> gpurun_out/r02_final_smoke.log
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/r02_final_tests.log 2>&1; echo rc=$? >> gpurun_out/r02_final_tests.log
( time timeout 1500 python bench.py --gpus 1 --steps 20 --warmup 5 ) > gpurun_out/r02_final_bench.log 2> gpurun_out/r02_final_bench.err
tail -n 2 gpurun_out/r02_final_smoke.log; tail -n 3 gpurun_out/r02_final_tests.log; tail -n 4 gpurun_out/r02_final_bench.err
