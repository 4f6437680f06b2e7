"""Where does an e2e step (bench.py step_e2e) spend its wall time?  Synchronises after each phase (so the sum is an
upper bound of the pipelined step)."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    from group_gan_gcn_gat_b200 import models as MD
    from group_gan_gcn_gat_b200.evaluate import evaluate_batch
    from group_gan_gcn_gat_b200.schedule import get_schedule
    torch.backends.cudnn.allow_tf32 = False
    dev = torch.device('cuda:0')
    data = bench.synth_batch(1 << 16, 1236)
    gen = MD.TrajectoryGenerator(obs_len=8, pred_len=12, embedding_dim=16, encoder_h_dim=32, decoder_h_dim=32, mlp_dim=64,
                                 noise_dim=(8,), noise_mix_type='global', pooling_type='pool_net',
                                 pool_every_timestep=False, bottleneck_dim=8, batch_norm=False, n_heads=1, alpha=0.2)
    gen.load_state_dict(bench.load_weights(), strict=True)
    gen = gen.to(dev).train()
    gen.pool_net.precision = 'bf16'
    host = {k: data[k].pin_memory() for k in ('obs_traj', 'obs_traj_rel', 'obs_traj_g', 'seq_start_end', 'pred_traj_gt')}
    out_host = torch.empty(2).pin_memory()
    sync = torch.cuda.synchronize

    def step(verbose):
        t = [time.perf_counter()]
        d = {k: host[k].to(dev, non_blocking=True) for k in ('obs_traj', 'obs_traj_rel', 'obs_traj_g', 'pred_traj_gt')}
        sync(); t.append(time.perf_counter())
        sse = host['seq_start_end'].clone()
        get_schedule(sse, dev)
        sync(); t.append(time.perf_counter())
        ade, fde = evaluate_batch(gen, d['obs_traj'], d['obs_traj_rel'], sse, d['obs_traj_g'], d['pred_traj_gt'], 20)
        t.append(time.perf_counter())          # host done issuing
        sync(); t.append(time.perf_counter())
        out_host.copy_(torch.stack([ade, fde]), non_blocking=True)
        sync(); t.append(time.perf_counter())
        if verbose:
            names = ['h2d', 'schedule', 'evaluate_batch host issue', 'evaluate_batch gpu drain', 'd2h']
            print('  '.join('%s %.2f ms' % (n, (b - a) * 1e3) for n, a, b in zip(names, t, t[1:])),
                  ' total %.2f ms' % ((t[-1] - t[0]) * 1e3))

    with torch.no_grad():
        for i in range(8):
            step(i >= 3)


if __name__ == '__main__':
    main()
