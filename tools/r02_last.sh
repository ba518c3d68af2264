set -x
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 2 --steps 8 --warmup 4 --no-predict > gpurun_out/r02_last2_bench2.log 2>&1; echo rc=$? >> gpurun_out/r02_last2_bench2.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29562 tests/multi_gpu_worker.py 24 > gpurun_out/r02_last2_worker.log 2>&1; echo rc=$? >> gpurun_out/r02_last2_worker.log
tail -n 2 gpurun_out/r02_last2_worker.log | cut -c1-200; tail -n 2 gpurun_out/r02_last2_bench2.log | cut -c1-400
