cd $GRAFT_REPO_ROOT
MGPU_ENV_ID=rware-small-4ag-v1 MGPU_ENVS=32768 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 tests/mgpu_check.py 2>&1 | tail -1 > gpurun_out/mgpu8_small4ag.json; cat gpurun_out/mgpu8_small4ag.json
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus 8 --steps 100 --warmup 10 --no-cpu-baseline --e2e-steps 2 2>&1 | tail -1 | cut -c1-200
