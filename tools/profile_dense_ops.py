"""Kernel breakdown (torch.profiler) of GATEncoder / GCNModule forward (+ backward) on dense-crowd scenes.
usage: python tools/profile_dense_ops.py [N] [bwd]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from group_gan_gcn_gat_b200 import modules as M  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
bwd = len(sys.argv) > 2
dev = torch.device('cuda:0')
s = max(1, 65536 // n)
tot = n * s
rng = np.random.RandomState(n)
lab = np.floor(rng.uniform(0, 1, tot) * max(1, n // 3)).astype(np.float32) + 1
lab[rng.uniform(0, 1, tot) < 0.1] = 0
lab = torch.from_numpy(lab).view(-1, 1).to(dev)
st = np.arange(s + 1) * n
sse = torch.from_numpy(np.stack([st[:-1], st[1:]], 1).astype(np.int64)).to(dev)
pos = torch.rand(tot, 2, device=dev)
for name, mod in (('GATEncoder', M.GATEncoder(None, 1, 0, 0.2)), ('GCNModule', M.GCNModule())):
    mod = mod.to(dev)
    x = torch.randn(tot, 40, device=dev, requires_grad=bwd)
    up = torch.randn(tot, 24, device=dev)

    def run():
        if bwd:
            mod.zero_grad(set_to_none=True)
            x.grad = None
            (mod(x, sse, pos, lab) * up).sum().backward()
        else:
            with torch.no_grad():
                mod(x, sse, pos, lab)
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA]) as prof:
        for _ in range(3):
            run()
        torch.cuda.synchronize()
    print('=====', name, 'N', n, 'scenes', s, 'bwd' if bwd else 'fwd', '(3 iterations)')
    print(prof.key_averages().table(sort_by='cuda_time_total', row_limit=14, max_name_column_width=70))
