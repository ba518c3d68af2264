set -x
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 3 --warmup 3 > gpurun_out/r02_c12_bench8.log 2>&1; echo rc=$? >> gpurun_out/r02_c12_bench8.log
tail -n 3 gpurun_out/r02_c12_bench8.log | cut -c1-600
