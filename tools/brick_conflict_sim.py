"""Host-only simulation of the brick kernels' shared-memory bank conflicts (wavefronts per 4-byte tap load, averaged over
three tiles x the 4 voxels of a thread) for the full affine G6 and the first 8 of the reference benchmark's random rotations:
best over the six warp layouts with TMA-compatible pitches (rows multiples of 16 bytes) against ANY row / plane pitch.
usage: python tools/brick_conflict_sim.py"""
import numpy as np, sys
sys.path.insert(0,'.')
from voltools_b200 import utils as U
n=512
c=np.divide(np.subtract((n,n,n),1),2,dtype=np.float32)
G6=U.transform_matrix(center=c, scale=(1.1,0.9,1.05), shear=(0.05,-0.03,0.02), rotation=(30,45,60), rotation_order='rzxz', translation=(5.5,-3.25,2.0))
rots=np.random.default_rng(1).uniform(-180,180,(100,3))[:8]
mats={'G6':G6}
for i,r in enumerate(rots):
    mats[f'rand{i}']=U.transform_matrix(rotation=tuple(r),rotation_order='sxyz',center=(256,256,256))
LAY=[(4,1),(3,2),(2,3),(3,1),(2,2),(4,0)]
VPT=4
def thread_pos(tid,layout):
    lwx,lwy=LAY[layout]; lwz=5-lwx-lwy
    lane=tid&31; warp=tid>>5
    lx=lane&((1<<lwx)-1); ly=(lane>>lwx)&((1<<lwy)-1); lz=lane>>(lwx+lwy)
    nwx=4-lwx; nwy=3-lwy
    wx=warp&((1<<nwx)-1); wy=(warp>>nwx)&((1<<nwy)-1); wz=warp>>(nwx+nwy)
    return (wx<<lwx)+lx,(wy<<lwy)+ly,(wz<<lwz)+lz
def wavefronts(addr):  # addr: (nwarps,32) word addresses
    tot=0
    for w in addr:
        u=np.unique(w)
        b=np.bincount(u%32,minlength=32)
        tot+=b.max()
    return tot/len(addr)
tids=np.arange(256)
for name,M in mats.items():
    M=np.asarray(M,np.float64)
    best_al=(9,None); best_free=(9,None)
    res={}
    for layout in range(6):
        pos=np.array([thread_pos(t,layout) for t in tids])
        tiles=[]
        for smp in range(3):
            a0_0=(n//3*(smp+1))//16*16; a1_0=(n//3*(smp+1))//8*8; a2_0=(n//3*(2-smp))//16*16
            for v in range(VPT):
                a=np.stack([a0_0+pos[:,2]*VPT+v, a1_0+pos[:,1], a2_0+pos[:,0], np.ones(256)],1)
                p=a@M[:3].T
                i=np.floor(p-0.5+0.5).astype(int)  # texel index floor(coord+0.5-0.5)
                tiles.append(i)
        ext=np.max([t.max(0)-t.min(0)+4 for t in tiles],0)  # brick extents z,y,x incl 4 taps
        # aligned: py multiple of 4 >= extx ; pz = py*bh, bh>=exty
        for py in range((ext[2]+3)//4*4, (ext[2]+3)//4*4+33, 4):
            for bh in range(ext[1], ext[1]+9):
                pz=py*bh
                wf=np.mean([wavefronts((t[:,0]*pz+t[:,1]*py+t[:,2]).reshape(8,32)) for t in tiles])
                if wf<best_al[0]: best_al=(wf,(layout,py,bh))
        for py in range(ext[2], ext[2]+12):
            for pz in range(py*ext[1], py*ext[1]+33):
                wf=np.mean([wavefronts((t[:,0]*pz+t[:,1]*py+t[:,2]).reshape(8,32)) for t in tiles])
                if wf<best_free[0]: best_free=(wf,(layout,py,pz))
    print(name,'ext',ext,'aligned best %.2f %s | free best %.2f %s'%(best_al[0],best_al[1],best_free[0],best_free[1]),flush=True)
