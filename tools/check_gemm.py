"""sgx_gemm against an fp64 product: relative error (to the largest |C|) and time for the shapes the library uses."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from group_gan_gcn_gat_b200 import ops  # noqa: E402

dev = 'cuda:0'
torch.manual_seed(0)
for (m, n, k) in [(1, 1, 1), (70, 33, 5), (513, 40, 72), (24, 32, 10000), (3, 2, 70001), (512, 32, 245081), (245081, 32, 512),
                  (245081, 512, 32), (65536, 72, 40)]:
    a = torch.randn(m, k, device=dev)
    b = torch.randn(k, n, device=dev)
    ref = a.double() @ b.double()
    out = ops.gemm(a, b)
    err = float((out.double() - ref).abs().max() / ref.abs().max())
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(5)]
    for s, e in ev:
        s.record()
        ops.gemm(a, b)
        e.record()
    torch.cuda.synchronize()
    ms = sorted(s.elapsed_time(e) for s, e in ev)[2]
    torch.backends.cuda.matmul.allow_tf32 = False
    err_t = float(((a @ b).double() - ref).abs().max() / ref.abs().max())
    print('gemm %-22s rel err %.2e (torch fp32: %.2e)  %.3f ms  %.1f TFLOP/s' % ((m, n, k), err, err_t, ms, 2e-9 * m * n * k / ms))
