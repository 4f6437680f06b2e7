"""Three GATEncoder + GCNModule forwards at the bench size (2^16 zara1-shaped scenes): the command the ncu captures of the
tcgen05 graph kernels are taken on."""
import os
import sys

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
x = torch.randn(n, 40, device=dev)
with torch.no_grad():
    for name, mod in (('gat', M.GATEncoder(None, 1, 0, 0.2)), ('gcn', M.GCNModule())):
        mod = mod.to(dev)
        for _ in range(3):
            out = mod(x, sse, pos, lab)
        torch.cuda.synchronize()
        print(name, 'fwd ok', float(out.abs().sum()))
