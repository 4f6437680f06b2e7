"""Short driver for ncu: three generator forwards on the bench batch (bf16 pooling)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from group_gan_gcn_gat_b200 import models as MD  # noqa: E402

torch.backends.cudnn.allow_tf32 = False
dev = torch.device('cuda:0')
data = bench.synth_batch(int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 16, 1236)
gen = MD.TrajectoryGenerator(obs_len=8, pred_len=12, embedding_dim=16, encoder_h_dim=32, decoder_h_dim=32, mlp_dim=64,
                             noise_dim=(8,), noise_mix_type='global', pooling_type='pool_net', pool_every_timestep=False,
                             bottleneck_dim=8, batch_norm=False, n_heads=1, alpha=0.2)
gen.load_state_dict(bench.load_weights(), strict=True)
gen = gen.to(dev).train()
gen.pool_net.precision = 'bf16'
d = {k: data[k].to(dev) for k in ('obs_traj', 'obs_traj_rel', 'obs_traj_g', 'seq_start_end')}
with torch.no_grad():
    for _ in range(3):
        gen(d['obs_traj'], d['obs_traj_rel'], d['seq_start_end'], d['obs_traj_g'])
torch.cuda.synchronize()
print('ok')
