"""Tiny end-to-end exercise of every kernel for compute-sanitizer memcheck (ragged indexing, atomics, TMEM)."""
import os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from group_gan_gcn_gat_b200 import models as MD, modules as M
torch.backends.cudnn.allow_tf32 = False
dev = 'cuda:0'
torch.manual_seed(0)
sizes = [1, 2, 5, 33, 3, 70, 2]
st = np.concatenate([[0], np.cumsum(sizes)])
sse = torch.tensor(np.stack([st[:-1], st[1:]], 1))
n = int(st[-1])
for prec, dims in (('fp32', (16, 32, 8)), ('bf16', (16, 32, 8)), ('bf16', (16, 48, 48)), ('fp32', (16, 48, 48))):
    m = M.PoolHiddenNet(embedding_dim=dims[0], h_dim=dims[1], mlp_dim=64, bottleneck_dim=dims[2], batch_norm=False,
                        precision=prec).to(dev)
    h = torch.randn(n, dims[1], device=dev, requires_grad=(prec == 'fp32'))
    pos = torch.rand(n, 2, device=dev) * 10
    out = m(h, sse, pos)
    if prec == 'fp32':
        out.sum().backward()
lab = torch.randint(0, 4, (n, 1), device=dev).float()
for mod in (M.GATEncoder(None, 1, 0, 0.2), M.GATEncoder(None, 2, 0, 0.2), M.GCNModule()):
    mod = mod.to(dev)
    x = torch.randn(n, 40, device=dev, requires_grad=True)
    mod(x, sse, pos, lab).sum().backward()
small = torch.tensor([[0, 3], [3, 5], [5, 12]])
x = torch.randn(12, 40, device=dev)
with torch.no_grad():
    M.GATEncoder(None, 1, 0, 0.2).to(dev)(x, small, pos[:12], lab[:12])          # fused warp-per-chunk path
gen = MD.TrajectoryGenerator(obs_len=8, pred_len=12, embedding_dim=16, encoder_h_dim=32, decoder_h_dim=32, mlp_dim=64,
                             noise_dim=(8,), noise_mix_type='global', pooling_type='pool_net', pool_every_timestep=False,
                             bottleneck_dim=8, batch_norm=False, n_heads=1).to(dev)
nb = 9000                                                                           # >= 8192: tensor-core LSTM path
s2 = torch.tensor(np.stack([np.arange(0, nb, 3), np.arange(3, nb + 3, 3)], 1))
with torch.no_grad():
    gen(torch.rand(8, nb, 2, device=dev), torch.randn(8, nb, 2, device=dev) * 0.3, s2, torch.ones(8, nb, 1, device=dev))
    gen(torch.rand(8, 12, 2, device=dev), torch.randn(8, 12, 2, device=dev) * 0.3, small, torch.ones(8, 12, 1, device=dev))
torch.cuda.synchronize()
print('sanitize run ok')
