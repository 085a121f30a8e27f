"""What plain streaming kernels reach on this box at the sizes of this path: torch.sum (pure read), memset and copy over
174 MB and 1 GB, a 512 MiB memset (L2 flush) before every call -> profiles/r02_stream_bw.json."""
import torch, json
dev='cuda'
flush=torch.empty(512<<20,dtype=torch.uint8,device=dev)
out={}
for mb in (174, 1024):
    x=torch.randn(mb*(1<<20)//4,device=dev)
    for _ in range(3): x.sum()
    ms=[]
    for _ in range(10):
        flush.zero_()
        a,b=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
        a.record(); x.sum(); b.record(); torch.cuda.synchronize(); ms.append(a.elapsed_time(b))
    ms.sort(); out['read_%dMB'%mb]={'ms':ms[5],'GBs':mb*1.048576/ms[5]}
    y=torch.empty_like(x)
    ms=[]
    for _ in range(10):
        flush.zero_()
        a,b=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
        a.record(); y.zero_(); b.record(); torch.cuda.synchronize(); ms.append(a.elapsed_time(b))
    ms.sort(); out['memset_%dMB'%mb]={'ms':ms[5],'GBs':mb*1.048576/ms[5]}
    ms=[]
    for _ in range(10):
        flush.zero_()
        a,b=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
        a.record(); y.copy_(x); b.record(); torch.cuda.synchronize(); ms.append(a.elapsed_time(b))
    ms.sort(); out['copy_%dMB'%mb]={'ms':ms[5],'GBs_read_plus_write':2*mb*1.048576/ms[5]}
    del x,y
print(json.dumps(out))
