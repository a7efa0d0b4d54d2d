import sys, numpy as np, torch
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import oracle_bind as oracle
import range_coder_rust_b200 as rcb
ctx=rcb.Context(0)
thr=oracle.zipf_thresholds(256,1.1)
for n,chunk in [(65536,65536),(65536*2,65536),(65536*32,65536),(65536*33,65536),(8<<20,65536)]:
    syms=oracle.generate(n,256,0x5EED0001,thr)
    d=torch.from_numpy(syms).cuda()
    model=ctx.model_from_counts(ctx.histogram(d,256))
    stream,offsets,nb=ctx.encode_chunks(d,chunk,model)
    nch=(n+chunk-1)//chunk
    status=torch.zeros(nch,dtype=torch.int32,device='cuda')
    try:
        out=ctx.decode_chunks(stream,offsets,n,chunk,model,status=status)
        err=None
    except Exception as e:
        err=str(e)[:60]
        out=ctx.decode_chunks(stream,offsets,n,chunk,model,status=status,sync=False); torch.cuda.synchronize()
    o=out.cpu().numpy(); st=status.cpu().numpy()
    bad=[]
    for i in range(nch):
        a=o[i*chunk:(i+1)*chunk]; b=syms[i*chunk:(i+1)*chunk]
        if not np.array_equal(a,b):
            bad.append((i,int(np.flatnonzero(a!=b)[0]),int(st[i])))
    print(n,chunk,'err',err,'nbad',len(bad),bad[:8], 'status nz', int((st!=0).sum()))
