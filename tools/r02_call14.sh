set -x
timeout 300 python -m pytest tests/test_gpu_gp.py -m gpu -q -x -k "blocked_cholesky or lml" > gpurun_out/r02_c14_tests.log 2>&1; echo rc=$? >> gpurun_out/r02_c14_tests.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 tests/multi_gpu_worker.py 24 > gpurun_out/r02_c14_worker.log 2>&1; echo rc=$? >> gpurun_out/r02_c14_worker.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus 2 --steps 2 --warmup 2 --no-predict > gpurun_out/r02_c14_bench2.log 2>&1; echo rc=$? >> gpurun_out/r02_c14_bench2.log
tail -n 5 gpurun_out/r02_c14_tests.log; tail -n 8 gpurun_out/r02_c14_worker.log | cut -c1-500; grep -o '"e2e": {[^}]*}[^}]*}[^}]*}' gpurun_out/r02_c14_bench2.log | cut -c1-900
