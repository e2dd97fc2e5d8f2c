# usage: bash gpu_job_mg.sh N   (run under gpurun --gpus N)
cd $GRAFT_REPO_ROOT
N=${1:-2}
mkdir -p gpurun_out
{
echo "== sysinfo"; cat /sys/fs/cgroup/cpu.max; nproc; nvidia-smi -L | head -8
if [ "$N" -ge 2 ]; then
echo "== multi-GPU gather test (2 ranks, slow consumer)"
timeout 900 python -m pytest tests/test_multigpu_gather.py tests/test_api_semantics.py -x -q -m gpu 2>&1 | tail -5
fi
echo "== e2e at N=$N (trace of rank 0)"
RWARE_B200_HOST_TRACE=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 profiles/hostexp/e2e_sweep.py rware-tiny-4ag-v1 65536 quick 2>&1 | grep -E "rw_hostpool|rw_step_host|env_steps" | awk 'NR%6==1 || /env_steps/' | tail -24
echo "== bench N=$N"
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 20 --warmup 5 2>&1 | tail -2
echo "== reference arm N=$N"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus $N --steps 1 --warmup 0 2>&1 | tail -1
} > gpurun_out/job_mg$N.log 2>&1
