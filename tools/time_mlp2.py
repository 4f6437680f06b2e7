"""CUDA-event time of the fused context MLP (ops.mlp2) at the bench size, L2 flushed.  Honours SGX_LIB=<variant.so>."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from group_gan_gcn_gat_b200 import modules as M, ops  # noqa: E402

dev = torch.device('cuda:0')
n = int(sys.argv[1]) if len(sys.argv) > 1 else 206477
torch.manual_seed(0)
mlp = M.make_mlp([40, 64, 24], batch_norm=False).to(dev)
xa, xb = torch.randn(n, 32, device=dev), torch.randn(n, 8, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
with torch.no_grad():
    ref = mlp(torch.cat([xa, xb], 1))
    ts = []
    for i in range(25):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        out = ops.mlp2(mlp, xa, xb)
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
ts = sorted(ts[5:])
print('lib=%s mlp2 %d rows: median %.1f us min %.1f us, max rel err vs torch %.2e' % (
    os.path.basename(os.environ.get('SGX_LIB', 'default')), n, ts[len(ts) // 2], ts[0],
    float((out - ref).abs().max() / ref.abs().max())))
