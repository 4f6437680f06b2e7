"""Short driver for ncu: a few launches of the bf16 pooling op on a dense-crowd batch and on the bench batch."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from group_gan_gcn_gat_b200 import modules as M  # noqa: E402

dev = 'cuda:0'
m = M.PoolHiddenNet(embedding_dim=16, h_dim=32, mlp_dim=64, bottleneck_dim=8, batch_norm=False,
                    precision=sys.argv[1] if len(sys.argv) > 1 else 'bf16').to(dev)
for sizes in ([1024] * 8, list(bench.synth_batch(1 << 16, 1236)['sizes'])):
    st = np.concatenate([[0], np.cumsum(sizes)])
    sse = torch.tensor(np.stack([st[:-1], st[1:]], 1))
    n = int(st[-1])
    h = torch.randn(n, 32, device=dev)
    pos = torch.rand(n, 2, device=dev) * 15
    with torch.no_grad():
        for _ in range(4):
            m(h, sse, pos)
    torch.cuda.synchronize()
print('ok')
