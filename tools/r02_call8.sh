set -x
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/multi_gpu_worker.py 24 > gpurun_out/r02_c8_worker.log 2>&1; echo rc=$? >> gpurun_out/r02_c8_worker.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 2 --warmup 1 --predict-structures 1280 > gpurun_out/r02_c8_bench2.log 2>&1; echo rc=$? >> gpurun_out/r02_c8_bench2.log
GPRB_BALANCE_INVERSE=0 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 2 --warmup 1 --no-predict > gpurun_out/r02_c8_bench2_nobal.log 2>&1
tail -n 12 gpurun_out/r02_c8_worker.log | cut -c1-400; tail -n 3 gpurun_out/r02_c8_bench2.log | cut -c1-3000
