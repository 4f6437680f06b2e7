"""SURVEY cfg 5: one adversarial training iteration (D step + G step with best_k = 20), data-parallel over scenes.

    python tools/train_step_dp.py                       # 1 GPU
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/train_step_dp.py

zara1-train-shaped synthetic batch (64 scenes per rank, scene sizes from the train histogram), fresh
kaiming weights (scripts/train.py:127-130), Adam, d_type global.  Prints one JSON line on rank 0 with the device
time per D / G step (max over ranks) and the all-reduced bytes.
"""
import json
import os
import sys
import time
from types import SimpleNamespace

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from group_gan_gcn_gat_b200 import models as MD, parallel  # noqa: E402

TRAIN_HIST = "2:399 3:269 4:237 5:187 6:141 7:107 8:64 9:76 10:34 11:23 12:23 13:29 14:15 15:19 16:17 17:22 18:32 19:23 20:31 21:35 22:40 23:36 24:35 25:37 26:22 27:16 28:19 29:28 30:29 31:24 32:31 33:29 34:30 35:16 36:20 37:20 38:21 39:22 40:7 41:5 42:6 43:9 44:4 45:7 46:6 47:2 48:4 49:3 50:1 51:5 52:2 53:2 57:1"


def synth(n_scenes, seed):
    rng = np.random.RandomState(seed)
    hist = dict((int(a), int(b)) for a, b in (kv.split(':') for kv in TRAIN_HIST.split()))
    sizes = rng.choice(list(hist), size=n_scenes, p=np.array(list(hist.values())) / sum(hist.values()))
    n = int(sizes.sum())
    st = np.concatenate([[0], np.cumsum(sizes)])
    disp = rng.normal(0, 0.3, size=(20, n, 2)).astype(np.float32)
    disp[0] = 0
    traj = rng.uniform(0, 15, size=(1, n, 2)).astype(np.float32) + np.cumsum(disp, 0)
    hi = np.maximum(1, np.repeat(sizes, sizes) // 3)
    lab = np.floor(rng.uniform(0, 1, n) * hi).astype(np.float32) + 1
    lab[rng.uniform(0, 1, n) < 0.13] = 0
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a))
    return dict(obs_traj=t(traj[:8]), pred_traj_gt=t(traj[8:]), obs_traj_rel=t(disp[:8]), pred_traj_gt_rel=t(disp[8:]),
                obs_traj_g=t(np.broadcast_to(lab[None, :, None], (8, n, 1)).copy()), loss_mask=torch.ones(n, 20),
                seq_start_end=t(np.stack([st[:-1], st[1:]], 1).astype(np.int64)))


def setup(dev, rank, world, scenes_per_rank=64):
    """-> (args, generator, discriminator, optimizer_g, optimizer_d, batch) for this rank; args.n_global = pedestrians of
    the global minibatch (no collective needed to find it)."""
    args = SimpleNamespace(obs_len=8, pred_len=12, best_k=20, l2_loss_weight=1.0, clipping_threshold_g=2.0,
                           clipping_threshold_d=0.0)
    torch.manual_seed(0)                                   # identical initial weights on every rank
    gen = MD.TrajectoryGenerator(obs_len=8, pred_len=12, embedding_dim=16, encoder_h_dim=32, decoder_h_dim=32, mlp_dim=64,
                                 noise_dim=(8,), noise_mix_type='global', pooling_type='pool_net',
                                 pool_every_timestep=False, bottleneck_dim=8, batch_norm=False, n_heads=1).to(dev)
    disc = MD.TrajectoryDiscriminator(obs_len=8, pred_len=12, embedding_dim=16, h_dim=48, mlp_dim=64, batch_norm=False,
                                      d_type='global').to(dev)
    for m in list(gen.modules()) + list(disc.modules()):
        if isinstance(m, torch.nn.Linear):
            torch.nn.init.kaiming_normal_(m.weight)
    with torch.no_grad():                                   # plain randn GCN weights are a passenger here
        for p in gen.gcn_module.parameters():
            p.mul_(0.1)
    opt_g = torch.optim.Adam(gen.parameters(), lr=1e-4)
    opt_d = torch.optim.Adam(disc.parameters(), lr=1e-3)
    data = synth(scenes_per_rank * world, 1239)
    tensors = {k: data[k] for k in ('obs_traj', 'pred_traj_gt', 'obs_traj_rel', 'pred_traj_gt_rel', 'obs_traj_g')}
    tensors['0:loss_mask'] = data['loss_mask']
    loc, sse, mine = parallel.shard_batch(tensors, data['seq_start_end'], world, rank)
    args.n_global = parallel.global_ped_count(data['seq_start_end'].numpy())
    batch = tuple(loc[k].to(dev) for k in ('obs_traj', 'pred_traj_gt', 'obs_traj_rel', 'pred_traj_gt_rel', 'obs_traj_g',
                                           'loss_mask')) + (sse.to(dev),)
    return args, gen, disc, opt_g, opt_d, batch


def main():
    rank, world = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1))
    local = int(os.environ.get('LOCAL_RANK', 0))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    torch.backends.cudnn.allow_tf32 = False
    args, gen, disc, opt_g, opt_d, batch = setup(dev, rank, world)
    times = {'d': [], 'g': []}
    for it in range(4):
        rng = parallel.make_label_rng(0, it)
        torch.manual_seed(100 + it * world + rank)          # noise differs per rank (different scenes)
        for name, fn, opt in (('d', parallel.discriminator_step, opt_d), ('g', parallel.generator_step, opt_g)):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            e0.record()
            losses = fn(args, batch, gen, disc, opt, label_rng=rng, n_global=args.n_global)
            e1.record()
            torch.cuda.synchronize()
            times[name].append(e0.elapsed_time(e1))
    if os.environ.get('SGX_PROFILE_STEP') and rank == 0:    # kernel / host breakdown of one G step (torch.profiler)
        rng = parallel.make_label_rng(0, 99)
        with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA,
                                                torch.profiler.ProfilerActivity.CPU]) as prof:
            t0 = time.perf_counter()
            parallel.generator_step(args, batch, gen, disc, opt_g, label_rng=rng, n_global=args.n_global)
            t1 = time.perf_counter()
            torch.cuda.synchronize()
            t2 = time.perf_counter()
        print('G step: host issue %.1f ms, drain %.1f ms' % ((t1 - t0) * 1e3, (t2 - t1) * 1e3), file=sys.stderr)
        print(prof.key_averages().table(sort_by='cuda_time_total', row_limit=45, max_name_column_width=60), file=sys.stderr)
    # parameters must stay bit-identical across ranks after the reduced updates
    flat = torch.cat([p.detach().reshape(-1) for p in gen.parameters()])
    chk = torch.stack([flat.double().sum(), flat.double().abs().sum()])
    if world > 1:
        lo, hi = chk.clone(), chk.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        in_sync = bool(torch.equal(lo, hi))
        t = torch.tensor([np.median(times['d'][1:]), np.median(times['g'][1:])], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        d_ms, g_ms = t.tolist()
    else:
        in_sync, d_ms, g_ms = True, float(np.median(times['d'][1:])), float(np.median(times['g'][1:]))
    if rank == 0:
        print(json.dumps({'workload': 'cfg5 adversarial step, 64 zara1-train-shaped scenes per rank, best_k=20',
                          'n_gpus': world, 'd_step_ms': d_ms, 'g_step_ms': g_ms, 'peds_rank0': int(batch[0].shape[1]),
                          'allreduce_bytes': {'G': sum(p.numel() for p in gen.parameters()) * 4,
                                              'D': sum(p.numel() for p in disc.parameters()) * 4},
                          'params_in_sync': in_sync, 'finite': bool(torch.isfinite(flat).all()),
                          'last_losses': {k: float(v) for k, v in losses.items()}}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
