"""CPU simulation of the pruning rule of csrc/fps_bucket.cu: how many of the 512 buckets (32 points each) of a KITTI-like
scene a round of D-FPS has to update when a bucket is skipped whenever its bounding box lies farther from the new sample
than its largest running distance.  Usage: python tools/fps_bucket_sim.py [bucket_size] [morton3|morton2|kd]
(numbers quoted in DESIGN.md section 3.1: ~6 affected buckets per round after the first thousand rounds)."""
import importlib, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
synth = importlib.import_module("3dssd_b200.synth")
pts = synth.kitti_like(1, 16384, seed=1000)[0,:,:3].astype(np.float32)
n=len(pts); m=4096
BS=int(sys.argv[1]) if len(sys.argv)>1 else 32
mode=sys.argv[2] if len(sys.argv)>2 else 'morton3'
lo=pts.min(0); hi=pts.max(0)
ext=hi-lo
def part1by2(v):
    v=v.astype(np.uint64)&0x3ff
    v=(v|(v<<16))&0x30000ff; v=(v|(v<<8))&0x300f00f; v=(v|(v<<4))&0x30c30c3; v=(v|(v<<2))&0x9249249
    return v
def part1by1(v):
    v=v.astype(np.uint64)&0xffff
    v=(v|(v<<8))&0x00ff00ff; v=(v|(v<<4))&0x0f0f0f0f; v=(v|(v<<2))&0x33333333; v=(v|(v<<1))&0x55555555
    return v
if mode=='morton3':
    cell=ext.max()/1024
    q=np.minimum(((pts-lo)/cell).astype(np.int64),1023)
    key=part1by2(q[:,0])|(part1by2(q[:,1])<<1)|(part1by2(q[:,2])<<2)
    order=np.argsort(key,kind='stable')
elif mode=='morton2':
    ax=np.argsort(-ext)[:2]
    cell=ext[ax].max()/32768
    q=np.minimum(((pts[:,ax]-lo[ax])/cell).astype(np.int64),32767)
    key=part1by1(q[:,0])|(part1by1(q[:,1])<<1)
    order=np.argsort(key,kind='stable')
elif mode=='kd':
    order=np.arange(n)
    def rec(idx,depth):
        if len(idx)<=BS: return [idx]
        e=pts[idx].max(0)-pts[idx].min(0); a=int(np.argmax(e))
        s=idx[np.argsort(pts[idx,a],kind='stable')]
        h=len(s)//2
        return rec(s[:h],depth+1)+rec(s[h:],depth+1)
    order=np.concatenate(rec(np.arange(n),0))
P=pts[order]
nb=n//BS
B=P.reshape(nb,BS,3)
bmin=B.min(1); bmax=B.max(1)
dist=np.full(n,1e10,np.float32)
cur=int(np.where(order==0)[0][0])
aff=[]; warpmax=[]; upd=[]
NW=32
for j in range(1,m):
    s=P[cur]
    d=np.maximum(0,np.maximum(bmin-s, s-bmax)); lb=(d*d).sum(1)
    bm=dist.reshape(nb,BS).max(1)
    a=lb<bm
    ids=np.where(a)[0]
    aff.append(len(ids))
    warpmax.append(np.bincount(ids%NW,minlength=NW).max())
    sub=B[ids]; dd=((sub-s)**2).sum(2).astype(np.float32)
    dv=dist.reshape(nb,BS)
    upd.append(int((dd<dv[ids]).sum()))
    dv[ids]=np.minimum(dv[ids],dd)
    cur=int(np.argmax(dist))
aff=np.array(aff); warpmax=np.array(warpmax); upd=np.array(upd)
print(mode,"BS",BS,"nb",nb)
for lo_,hi_ in [(0,16),(16,64),(64,256),(256,1024),(1024,4095)]:
    print("rounds %d-%d: mean affected buckets %.1f  max-per-warp mean %.2f  p99 %.0f  actual point updates %.1f"%(lo_,hi_,aff[lo_:hi_].mean(),warpmax[lo_:hi_].mean(),np.percentile(warpmax[lo_:hi_],99),upd[lo_:hi_].mean()))
print("total bucket updates",aff.sum(),"sum of per-round max-per-warp",warpmax.sum())
