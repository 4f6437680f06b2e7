"""Two PoolHiddenNet forward + backward passes at the bench size (2^16 zara1-shaped scenes): the command the launch list /
ncu capture of the pooling backward is taken on."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from group_gan_gcn_gat_b200 import modules as M  # noqa: E402

dev = torch.device('cuda:0')
data = bench.synth_batch(1 << 16, 1237, 'sgan_gat')
sse = data['seq_start_end'].to(dev)
n = int(sse[-1, 1])
pos = data['obs_traj'][-1].to(dev)
torch.manual_seed(0)
pool = M.PoolHiddenNet(embedding_dim=16, h_dim=32, mlp_dim=64, bottleneck_dim=8, batch_norm=False).to(dev)
h = torch.randn(n, 32, device=dev, requires_grad=True)
for _ in range(2):
    out = pool(h, sse, pos)
    go = torch.randn_like(out)
    g = torch.autograd.grad(out, [h] + list(pool.parameters()), go)
torch.cuda.synchronize()
print('pool fwd+bwd ok', n, float(g[0].abs().sum()))
