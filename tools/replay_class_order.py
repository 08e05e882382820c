#!/usr/bin/env python
"""Offline replay of the candidate-record stream of the centred scoring pass (score_tc2.cu) on dumped C4 factors
(gpurun_out/conv_c4.npz from tools/dump_converged.py): how many records (items that beat the row's running n-th best) a
row produces per sweep under different scan orders of the item classes, with and without threshold lag.

    python tools/replay_class_order.py last|early

Printed per order: records per row (mean, p99) with a lag-free threshold (`ideal`), with a threshold that is one 32-column
block old, and with a private bootstrap on the first 32 columns.  Result on the 25-iteration factors: rising-norm order
(r1 / early r2) 165 ideal, falling sampled density (what class_scan_kernel does) 76 ideal / 130 with lag.
"""
import numpy as np

d = np.load('gpurun_out/conv_c4.npz')
import sys
tag=sys.argv[1]
V=d['v_'+tag].astype(np.float64); U=d['u_'+tag][:300].astype(np.float64)
ni=V.shape[0]
ip=d['indptr_s']; ix=d['indices_s']
vbar=V.mean(0); W=V-vbar
wn=np.linalg.norm(W,axis=1)
cls=np.clip(np.floor(np.log2(wn.max()/np.maximum(wn,1e-30))).astype(int),0,7)
S=U@W.T
def sim(order, blk=32, lag=0, bootn=0):
    recs=[]
    for r in range(U.shape[0]):
        s=S[r][order]
        seen=np.zeros(ni,bool); sl=ix[ip[r]:ip[r+1]]; sl=sl[sl<ni]; seen[sl]=True; seen=seen[order]
        su=np.where(seen,-np.inf,s)
        nb=(ni+blk-1)//blk; pad=nb*blk-ni
        sp=np.concatenate([s,np.full(pad,-np.inf)]).reshape(nb,blk)
        sup=np.concatenate([su,np.full(pad,-np.inf)]).reshape(nb,blk)
        top=np.full(10,-np.inf); th=np.empty(nb)
        for b in range(nb):
            top=np.partition(np.concatenate([top,sup[b]]),-10)[-10:]
            th[b]=top.min()
        use=np.full(nb,-np.inf)
        if nb>1+lag: use[1+lag:]=th[:nb-1-lag]
        if bootn: use[:1+lag]=np.sort(sup[0][:bootn])[-10]
        recs.append((sp>use[:,None]).sum())
    return round(np.mean(recs),1), round(np.percentile(recs,99),1)
# sample users (different from test users) -> top10 histogram
Us=d['u_'+tag][4000:4064].astype(np.float64); Ss=Us@W.T
hist=np.zeros(8)
for r in range(64):
    t=np.argpartition(Ss[r],-10)[-10:]; hist+=np.bincount(cls[t],minlength=8)
size=np.bincount(cls,minlength=8)
print(tag,'hist',hist,'size',size)
def order_from(rank):
    inv=np.empty(8,int); inv[np.array(rank)]=np.arange(8)
    return np.lexsort((np.arange(ni),inv[cls]))
cands={'current(asc |w|)':[7,6,5,4,3,2,1,0],'desc |w|':[0,1,2,3,4,5,6,7],
 'by count':list(np.argsort(-hist,kind='stable')),'by density':list(np.argsort(-(hist/np.maximum(size,1)),kind='stable'))}
for name,rk in cands.items():
    o=order_from(rk)
    print(tag,name,rk,'ideal',sim(o,1,0),'blk32lag1',sim(o,32,1),'boot',sim(o,32,1,32))
