set -x
cd $GRAFT_REPO_ROOT
python profiles/membw_probe.py 2>&1 | tail -1
ncu --set full --clock-control none --import-source on -k regex:rw_main_kernel -s 1 -c 1 -o gpurun_out/prof_obs_r1 python -c "
import torch, rware_b200
v = rware_b200.make_vec('rware-tiny-4ag-v1', 65536)
v.reset(seed=0)
for i in range(3): v.observations()
torch.cuda.synchronize()
" > gpurun_out/ncu_obs.log 2>&1
tail -2 gpurun_out/ncu_obs.log
