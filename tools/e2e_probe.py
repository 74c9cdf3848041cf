import sys, time, numpy as np, torch
sys.path.insert(0, "/root/repo")
from vector_db_id_compression_b200.capi import Context
from vector_db_id_compression_b200 import workloads as W, capi
N=int(1e9); dev=torch.device("cuda:0")
sizes=W.zipf_sizes(N,65536,1.0); offsets,ids=W.random_partition_lists(N,sizes,1234,dev)
hin=torch.empty(N,dtype=torch.int64,pin_memory=True); hin.copy_(ids); hout=torch.empty(N,dtype=torch.int64,pin_memory=True)
dbuf=torch.empty_like(ids); torch.cuda.synchronize()
for name,fn in [("h2d",lambda: dbuf.copy_(hin,non_blocking=True)),("d2h",lambda: hout.copy_(dbuf,non_blocking=True))]:
    for _ in range(2):
        t=time.perf_counter(); fn(); torch.cuda.synchronize(); dt=time.perf_counter()-t
    print(name, "%.1f ms %.1f GB/s"%(dt*1e3, 8*N/dt/1e9))
ctx=Context(0, stream=torch.cuda.current_stream().cuda_stream); ctx.set_timing(True)
h=hin.numpy(); ho=hout.numpy()
for rep in range(3):
    t0=time.perf_counter(); blob=ctx.roc_encode(offsets,h,sorted_ids=True); t1=time.perf_counter()
    kb=dict(ctx.last_kernel_breakdown())
    p,mem=capi._ptr(ho); off=np.zeros(blob.nlist+1,np.uint64)
    capi._check(ctx._l.idc_roc_decode(ctx._h, blob._h, None, blob.nlist, p, 8, mem, off.ctypes.data)); t2=time.perf_counter()
    kd=dict(ctx.last_kernel_breakdown())
    blob.free(); t3=time.perf_counter()
    print("encode(host) %.1f ms  decode(host) %.1f ms  free %.1f ms | enc kernels %.1f dec kernels %.1f"%((t1-t0)*1e3,(t2-t1)*1e3,(t3-t2)*1e3,sum(kb.values()),sum(kd.values())))
