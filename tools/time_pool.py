"""Kernel-only timing of the pooling op (library-recorded events around the dominant kernel)."""
import ctypes
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from group_gan_gcn_gat_b200 import _lib, modules as M  # noqa: E402


def time_case(sizes, dims, precision, reps=10):
    dev = 'cuda:0'
    e, h_dim, b = dims
    m = M.PoolHiddenNet(embedding_dim=e, h_dim=h_dim, mlp_dim=64, bottleneck_dim=b, batch_norm=False, precision=precision).to(dev)
    st = np.concatenate([[0], np.cumsum(sizes)])
    sse = torch.tensor(np.stack([st[:-1], st[1:]], 1))
    n = int(st[-1])
    h = torch.randn(n, h_dim, device=dev)
    pos = torch.rand(n, 2, device=dev) * 15
    L = _lib.lib()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); e1.record(); torch.cuda.synchronize()
    L.sgx_profile_events(ctypes.c_void_p(e0.cuda_event), ctypes.c_void_p(e1.cuda_event))
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    ts = []
    with torch.no_grad():
        for r in range(reps + 3):
            flush.fill_(r)
            m(h, sse, pos)
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
    L.sgx_profile_events(None, None)
    ts = sorted(ts[3:])
    pairs = int((np.asarray(sizes, dtype=np.int64) ** 2).sum())
    flops = pairs * (4 * e + 2 * (e + h_dim) * 512 + 2 * 512 * b)
    med = ts[len(ts) // 2]
    print('%-28s %-9s pairs=%-10d kernel_ms med=%.4f min=%.4f  as-written TFLOP/s=%.1f  Gpairs/s=%.2f' % (
        'N=%s x%d' % (sizes[0], len(sizes)), precision, pairs, med, ts[0], flops / med / 1e9, pairs / med / 1e6))


if __name__ == '__main__':
    G, D = (16, 32, 8), (16, 48, 48)
    precs = sys.argv[1:] or ('fp32-simt', 'tc32', 'bf16')
    for prec in precs:
        time_case([1024] * 8, G, prec)
        time_case([256] * 64, G, prec)
        time_case([64] * 512, G, prec)
        z = bench.synth_batch(1 << 16, 1236)['sizes']
        time_case(list(z), G, prec)
        if prec != 'tc32':
            time_case([1024] * 4, D, prec)
