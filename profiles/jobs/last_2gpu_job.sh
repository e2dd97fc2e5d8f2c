cd $GRAFT_REPO_ROOT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517"
timeout 300 $TR tests/mgpu_check.py 2>&1 | tail -3
timeout 300 $TR bench.py --gpus 2 --steps 200 --warmup 20 --no-cpu-baseline --e2e-steps 2 2>/dev/null | tail -1 > gpurun_out/bench_2gpu.json; cut -c1-200 gpurun_out/bench_2gpu.json
timeout 300 $TR bench.py --gpus 2 --gather nccl --steps 100 --warmup 10 --no-cpu-baseline --e2e-steps 2 2>/dev/null | tail -1 > gpurun_out/bench_2gpu_nccl.json; cut -c1-200 gpurun_out/bench_2gpu_nccl.json
timeout 300 $TR bench.py --gpus 2 --gather fused --steps 100 --warmup 10 --no-cpu-baseline --e2e-steps 2 2>/dev/null | tail -1 > gpurun_out/bench_2gpu_fused.json; cut -c1-200 gpurun_out/bench_2gpu_fused.json
