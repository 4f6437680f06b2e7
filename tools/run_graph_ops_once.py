"""One GATEncoder / GCNModule forward + backward at the bench size (2^16 zara1-shaped scenes) and one dense-crowd
GATEncoder forward (64 scenes of 1024): the command the ncu captures of the single-launch backward kernels and of the
warp-per-row scene kernels are taken on (profiles/r02_graph_kernels_ncu_full_raw.csv)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from group_gan_gcn_gat_b200 import modules as M  # noqa: E402

dev = torch.device('cuda:0')
data = bench.synth_batch(1 << 16, 1237)
sse = data['seq_start_end'].to(dev)
n = int(sse[-1, 1])
lab, pos = data['obs_traj_g'][-1].to(dev), data['obs_traj'][-1].to(dev)
torch.manual_seed(0)
for name, mod in (('gat', M.GATEncoder(None, 1, 0, 0.2)), ('gcn', M.GCNModule())):
    mod = mod.to(dev)
    with torch.no_grad():
        for p in mod.parameters():
            if name == 'gcn' and p.dim() == 2 and tuple(p.shape) != (24, 32):
                p.mul_(0.15)
    x = torch.randn(n, 40, device=dev, requires_grad=True)
    up = torch.randn(n, 24, device=dev)
    for _ in range(2):
        mod.zero_grad(set_to_none=True)
        x.grad = None
        (mod(x, sse, pos, lab) * up).sum().backward()
    torch.cuda.synchronize()
    print(name, 'fwd+bwd ok', float(x.grad.abs().sum()))
nd, s = 1024, 64
rng = np.random.RandomState(1)
labd = np.floor(rng.uniform(0, 1, nd * s) * (nd // 3)).astype(np.float32) + 1
labd[rng.uniform(0, 1, nd * s) < 0.1] = 0
st = np.arange(s + 1) * nd
ssed = torch.from_numpy(np.stack([st[:-1], st[1:]], 1).astype(np.int64)).to(dev)
gat = M.GATEncoder(None, 1, 0, 0.2).to(dev)
with torch.no_grad():
    for _ in range(2):
        out = gat(torch.randn(nd * s, 40, device=dev), ssed, torch.rand(nd * s, 2, device=dev),
                  torch.from_numpy(labd).view(-1, 1).to(dev))
torch.cuda.synchronize()
print('dense gat ok', float(out.abs().sum()))
