import ctypes, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from group_gan_gcn_gat_b200 import _lib
from tools.time_pool import time_case
DIMS = (16, 48, 48) if len(sys.argv) > 1 and sys.argv[1] == 'd' else (16, 32, 8)
PREC = sys.argv[2] if len(sys.argv) > 2 else 'bf16'     # tc32: G1 (x_full, buf_free), G2 (-, -, d2_free, h_ready),
#                                                         ROW0 as below, EPI (d1_full wait, convert cycles)
L = _lib.lib()
buf = torch.zeros(40, dtype=torch.int64, device='cuda')
h = ctypes.CDLL(_lib.LIB_PATH)
h.sgx_debug_tc_stats.argtypes = [ctypes.c_void_p]
h.sgx_debug_tc_stats(buf.data_ptr())
names = ['G1 issuer (x_full, d1_free)', 'G2 issuer (-, -, d2_free, h_ready)', 'ROW0 (wait x_free, wait d2_full, prefetch cycles, finalize cycles)', 'EPI0 (d1_full, h_free)', 'EPI1 (d1_full, h_free)']
for label, sizes in (('dense N=1024 x8', [1024] * 8), ('zara-shaped 65536 scenes', list(bench.synth_batch(1 << 16, 1236)['sizes']))):
    time_case(sizes, DIMS, PREC, reps=3)
    torch.cuda.synchronize()
    s = buf.cpu().view(5, 8)
    for r in range(5):
        n = max(1, int(s[r, 5]))
        print(label, '|', names[r], 'per-tile wait cycles:', [int(v) // n for v in s[r, :4]], 'total/tile:', int(s[r, 4]) // n, 'tiles', n)
