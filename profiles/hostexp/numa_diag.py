"""Where do the pages of a buffer from rw_host_alloc_near end up?  Prints the NUMA node of every 2 MB stripe."""
import ctypes, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import rware_b200
print(open("/proc/self/status").read().split("Mems_allowed_list:")[1].split()[0], "mems allowed;",
      "THP:", open("/sys/kernel/mm/transparent_hugepage/enabled").read().strip())
v = rware_b200.make_vec("rware-tiny-4ag-v1", 65536)
be = v.backend
t = be.host_alloc((65536, 4, 71), torch.float32)
base = t.data_ptr()
nodes = [be.lib.rw_host_numa_node(ctypes.c_void_p(base + off)) for off in range(0, t.numel() * 4, 2 << 20)]
print("gpu node", be.gpu_numa_node(), "stripes:", "".join(str(x) if x >= 0 else "?" for x in nodes))
nodes4k = [be.lib.rw_host_numa_node(ctypes.c_void_p(base + off)) for off in range(0, 8 << 20, 4096 * 64)]
print("first 8 MB every 256 KB:", "".join(str(x) if x >= 0 else "?" for x in nodes4k))
g, th = be.host_stream_probe(t, 3)
print("probe GB/s", g, "threads", th)
