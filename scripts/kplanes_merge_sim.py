import torch, numpy as np, sys
sys.path.insert(0,'/root/repo')
from tinynerf_b200 import synthetic
from oracle import ref_port as rp
torch.manual_seed(0)
o,d = synthetic.blender_rays(10240, seed=5)
grid = synthetic.analytic_grid(128, seed=3)
aabb = torch.tensor([[-1.5]*3,[1.5]*3])
noise = torch.rand(10240,256)
thr = 0.01
p, info, _ = rp.ray_provider(o,d,grid,thr,scene='aabb',n_samples=256,aabb=aabb,noise=noise)
x = p[:, :3].numpy().astype(np.float64)
N = len(x); print('samples',N,'rays with samples',(info[:,1]>0).sum().item())
ray_id = np.repeat(np.arange(len(info)), info[:,1].numpy())
for res in (128,256,512):
    tot_direct=0; 
    out={}
    for G in (4,8,16,32):
        out[G]=0
    runs_cell=0; runs_line=0
    for (a,b) in ((0,1),(0,2),(1,2)):
        iu = (x[:,a]+1)*0.5*(res-1); iv=(x[:,b]+1)*0.5*(res-1)
        x0=np.floor(iu).astype(np.int64); y0=np.floor(iv).astype(np.int64)
        cell = y0*res+x0
        # corner line ids
        corners = np.stack([cell, cell+1, cell+res, cell+res+1],1)  # [N,4]
        tot_direct += 4*N
        for G in (4,8,16,32):
            ng = N//G
            c = corners[:ng*G].reshape(ng, G*4)
            c.sort(axis=1)
            distinct = 1 + (np.diff(c,axis=1)!=0).sum(1)
            out[G]+= distinct.sum()
        # serial run-merge: consecutive samples same cell (within chunk of 16)
        same = (cell[1:]==cell[:-1])
        runs_cell += N - same.sum()
        # pair-line (two x-adjacent corners = 256B contiguous) distinct rows: row id (y, x0)
    print(res, {G: round(out[G]/tot_direct,3) for G in out}, 'cell-runs frac', round(runs_cell/(3*N),3))
