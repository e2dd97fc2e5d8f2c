cd $GRAFT_REPO_ROOT
ncu --set full --clock-control none --import-source on -k regex:rw_main_kernel -s 4 -c 1 -o gpurun_out/prof_step_r1h python profiles/prof_step.py > gpurun_out/ncu_h.log 2>&1
tail -2 gpurun_out/ncu_h.log
python - <<'PY'
import sys, os, json, numpy as np, torch
sys.path.insert(0, os.getcwd())
import rware_b200
# per-CTA end-time distribution (stragglers)
E=65536
vec = rware_b200.make_vec("rware-tiny-4ag-v1", E, auto_reset=True); vec.reset(seed=0)
acts = torch.empty((E, 4), dtype=torch.int32, device="cuda"); vec.backend.random_actions(acts, 1, 0)
for _ in range(20): vec.step(acts)
buf = torch.zeros((256*32*8,), dtype=torch.int64, device="cuda"); vec.backend.set_clock_buffer(buf)
torch.cuda.synchronize(); vec.step(acts); torch.cuda.synchronize()
t = buf.cpu().numpy().reshape(256,32,8).astype(np.float64); used = t[:,:,0]>0
t0 = t[:,:,0][used].min()
start = np.array([t[c,:,0][used[c]].min() for c in range(147)]) - t0
end = np.array([t[c,:,6][used[c]].max() for c in range(147)]) - t0
dur = end-start
print("cta start us: pct", np.percentile(start,[0,50,90,100])/1e3)
print("cta end   us: pct", np.percentile(end,[0,50,90,100])/1e3)
print("cta dur   us: pct", np.percentile(dur,[0,50,90,100])/1e3)
lg = np.array([ (t[c,0::2,3][used[c,0::2]] - t[c,0::2,2][used[c,0::2]]).mean() for c in range(147)])/1e3
print("logic us per cta pct", np.percentile(lg,[0,50,90,100]))
PY
