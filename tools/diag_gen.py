import sys; sys.path.insert(0,'.')
import torch, numpy as np
from apss_b200 import synth
kw=dict(N=6000, D=1<<14, nnz_mean=40, seed=20260103)
res={}
for dev in ('cpu','cuda'):
    g=synth.Generator(kw['N'],kw['D'],kw['nnz_mean'],seed=kw['seed'],device=dev)
    indptr,dims,tf,is_dup=g.structure(0,kw['N'])
    df=torch.bincount(dims.long(),minlength=kw['D'])
    idf=synth.idf_from_df(df,kw['N'])
    n=indptr.numel()-1
    cnt=indptr[1:]-indptr[:-1]
    row_of=torch.repeat_interleave(torch.arange(n,device=dev,dtype=torch.int64),cnt)
    u=synth._u53(synth._hash3(g.seed,synth._S_JIT,row_of,dims.long()))
    uf=u.to(torch.float64)
    m=uf*(0.2/float(1<<53))
    jit=m+0.9
    jit=torch.where(is_dup[row_of],jit,torch.ones_like(jit))
    a=tf.to(torch.float64)*idf.to(dev)[dims.long()]
    val=a*jit
    sq=val*val
    out=g.finalize(0,indptr,dims,tf,is_dup,idf)
    res[dev]=dict(u=u.cpu(),uf=uf.cpu(),m=m.cpu(),jit=jit.cpu(),a=a.cpu(),val=val.cpu(),sq=sq.cpu(),final=out.values.cpu(),idf=idf)
for k in res['cpu']:
    x,y=res['cpu'][k],res['cuda'][k]
    print(k, torch.equal(x,y), int((x!=y).sum()))
# norms
x,y=res['cpu']['final'],res['cuda']['final']
bad=(x!=y).nonzero().flatten()[:5]
print(bad, x[bad], y[bad])
print('---- stage 2')
res2={}
for dev in ('cpu','cuda'):
    g=synth.Generator(kw['N'],kw['D'],kw['nnz_mean'],seed=kw['seed'],device=dev)
    indptr,dims,tf,is_dup=g.structure(0,kw['N'])
    n=indptr.numel()-1
    cnt=indptr[1:]-indptr[:-1]
    row_of=torch.repeat_interleave(torch.arange(n,device=dev,dtype=torch.int64),cnt)
    sq=res[dev]['sq'].to(dev); val=res[dev]['val'].to(dev)
    pos=torch.arange(dims.numel(),device=dev,dtype=torch.int64)-indptr[:-1][row_of]
    acc=torch.zeros(n,dtype=torch.float64,device=dev)
    order=torch.argsort(pos*n+row_of)
    sqc,rowc=sq[order],row_of[order]
    colcnt=torch.bincount(pos,minlength=int(cnt.max()))
    ends=torch.cumsum(colcnt,0).tolist(); st=0
    for e in ends:
        acc[rowc[st:e]]=acc[rowc[st:e]]+sqc[st:e]; st=e
    nrm=torch.sqrt(acc)
    fin=val/nrm[row_of]
    res2[dev]=dict(order=order.cpu(),acc=acc.cpu(),nrm=nrm.cpu(),fin=fin.cpu(),nr=nrm[row_of].cpu())
for k in res2['cpu']:
    x,y=res2['cpu'][k],res2['cuda'][k]
    print(k, torch.equal(x,y), int((x!=y).sum()))
x=res2['cpu']['nr']; v=res['cpu']['val']
print('div on cpu of same inputs vs cuda:', int(((v/x)!=(v.cuda()/x.cuda()).cpu()).sum()))
a=res2['cpu']['acc']; print('sqrt cpu vs cuda on same inputs:', int((torch.sqrt(a)!=torch.sqrt(a.cuda()).cpu()).sum()))
