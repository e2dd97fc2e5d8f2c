cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_known_answers.py -m gpu -q 2>&1 | tail -2
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 tests/mgpu_check.py 2>&1 | tail -1
MGPU_ENVS=1000 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 tests/mgpu_check.py 2>&1 | tail -1
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29543 bench.py --gpus 2 --steps 50 --warmup 5 --no-cpu-baseline --e2e-steps 2 --gather fused 2>&1 | tail -1 | cut -c1-420
