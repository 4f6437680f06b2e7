"""One launch of every kernel added late in round 2, at the bench size: tcgen05 GATEncoder / GCNModule / context-MLP forwards,
the scene-owned pooling backward and the tensor-core GEMMs behind it (the command of profiles/r02_late_kernels_ncu_full_raw.csv)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from group_gan_gcn_gat_b200 import modules as M, ops  # noqa: E402

dev = torch.device('cuda:0')
data = bench.synth_batch(1 << 16, 1237, 'sgan_gat')
sse = data['seq_start_end'].to(dev)
n = int(sse[-1, 1])
lab, pos = data['obs_traj_g'][-1].to(dev), data['obs_traj'][-1].to(dev)
torch.manual_seed(0)
x = torch.randn(n, 40, device=dev)
gat, gcn = M.GATEncoder(None, 1, 0, 0.2).to(dev), M.GCNModule().to(dev)
mlp = M.make_mlp([40, 64, 24], batch_norm=False).to(dev)
pool = M.PoolHiddenNet(embedding_dim=16, h_dim=32, mlp_dim=64, bottleneck_dim=8, batch_norm=False).to(dev)
for _ in range(2):
    with torch.no_grad():
        a = gat(x, sse, pos, lab)
        b = gcn(x, sse, pos, lab)
        c = ops.mlp2(mlp, x[:, :32].contiguous(), x[:, 32:].contiguous())
    h = torch.randn(n, 32, device=dev, requires_grad=True)
    out = pool(h, sse, pos)
    g = torch.autograd.grad(out, [h] + list(pool.parameters()), torch.randn_like(out))
torch.cuda.synchronize()
print('ok', float(a.abs().sum()), float(b.abs().sum()), float(c.abs().sum()), float(g[0].abs().sum()))
