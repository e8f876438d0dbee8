import sys, numpy as np, torch, time
sys.path.insert(0,'/root/repo')
from cgraphs_b200.mdhbond import _native
ctx=_native.Context(0)
rng=np.random.default_rng(0)
for world,cap in [(2,2048),(4,2048),(8,2048)]:
    g=np.zeros((world,cap+1),dtype=np.int64)
    base=np.sort(rng.choice(1<<40,size=1900,replace=False))
    for w in range(world):
        k=np.sort(rng.choice(base,size=1507,replace=False)); g[w,0]=len(k); g[w,1:1+len(k)]=k; g[w,1+len(k):]=-1
    gt=torch.from_numpy(g).cuda(); out=torch.empty(world*cap,dtype=torch.int64,device='cuda'); info=torch.empty(2,dtype=torch.int64,device='cuda')
    st=torch.cuda.Stream(); ctx.set_option("compute_stream", st.cuda_stream)
    with torch.cuda.stream(st):
        for _ in range(3): ctx.merge_keys_device(gt.data_ptr(),world,cap,out.data_ptr(),info.data_ptr())
        e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
        e0.record(st)
        for _ in range(20): ctx.merge_keys_device(gt.data_ptr(),world,cap,out.data_ptr(),info.data_ptr())
        e1.record(st); torch.cuda.synchronize()
    print(world,cap,"merge us",1000*e0.elapsed_time(e1)/20, info.tolist())
