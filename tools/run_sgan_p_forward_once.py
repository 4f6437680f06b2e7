"""Short driver for ncu: three SGAN-P generator forwards on the default bench batch (fp32 contract pooling), the kernels of
the chain of DESIGN.md 4.10.

    ncu --set full --clock-control none --import-source on -k regex:"lstm_tc_kernel|pool_tc32_kernel" -s 6 -c 3 \
        -o gpurun_out/r02_final_chain python tools/run_sgan_p_forward_once.py
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

dev = torch.device('cuda:0')
torch.cuda.set_device(dev)
gen = bench.build_generator('sgan_p', dev)
data = bench.synth_batch(int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 16, 0, 'sgan_p')
d = {k: data[k].to(dev) for k in ('obs_traj', 'obs_traj_rel', 'obs_traj_g', 'seq_start_end')}
with torch.no_grad():
    for _ in range(3):
        gen(d['obs_traj'], d['obs_traj_rel'], d['seq_start_end'], d['obs_traj_g'])
torch.cuda.synchronize()
print('ok')
