import ctypes, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from group_gan_gcn_gat_b200 import _lib, models as MD
L = _lib.lib()
buf = torch.zeros(24, dtype=torch.int64, device='cuda')
h = ctypes.CDLL(_lib.LIB_PATH)
h.sgx_debug_lstm_stats.argtypes = [ctypes.c_void_p]
h.sgx_debug_lstm_stats(buf.data_ptr())
enc = MD.Encoder(embedding_dim=16, h_dim=32, mlp_dim=64).cuda()
x = torch.randn(8, 245081, 2, device='cuda') * 0.3
with torch.no_grad():
    for _ in range(3):
        enc(x)
torch.cuda.synchronize()
s = buf.cpu().view(3, 8)
names = ['issuer (wait a_ready slot0, slot1)', 'epi slot0 (wait g_full, gates math, write_a_rows, fence+arrive)', 'epi slot1']
for r in range(3):
    n = max(1, int(s[r, 5]))
    print(names[r], [int(v) // n for v in s[r, :4]], 'total per round-step', int(s[r, 4]) // n, 'round-steps', n)
