set -x
cd $GRAFT_REPO_ROOT
nvidia-smi -L
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 100 --warmup 10 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench_2gpu.json; cut -c1-700 gpurun_out/bench_2gpu.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 100 --warmup 10 --no-cpu-baseline --gather 2>&1 | tail -1 > gpurun_out/bench_2gpu_gather.json; cut -c1-700 gpurun_out/bench_2gpu_gather.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus 2 --steps 1 --warmup 0 --ref-seconds 3 2>&1 | tail -1 | cut -c1-300
