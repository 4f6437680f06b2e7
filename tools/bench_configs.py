"""Measurements of the other BASELINE.json configs next to bench.py's headline:
  cfg 3  GCNModule / GATEncoder op-level forward + backward at S = 2^16 zara1-shaped scenes
  cfg 4  dense crowds: N in {64..1024}, pool_every_timestep = 1, pred 12 (13 pooling calls per forward)
One JSON line per measurement (CUDA events, median of reps, L2 flushed between reps)."""
import json
import os
import statistics
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from group_gan_gcn_gat_b200 import models as MD, modules as M  # noqa: E402

dev = torch.device('cuda:0')
torch.backends.cudnn.allow_tf32 = False
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json'))) if os.path.exists(os.path.join(ROOT, 'MEASURED_PEAKS.json')) else {}


def timed(fn, reps=7, warm=3):
    ts = []
    for i in range(reps + warm):
        flush.fill_(i)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return statistics.median(ts[warm:])


def sse_of(sizes):
    st = np.concatenate([[0], np.cumsum(sizes)])
    return torch.from_numpy(np.stack([st[:-1], st[1:]], 1).astype(np.int64))


def cfg3():
    data = bench.synth_batch(1 << 16, 1237)
    sse = data['seq_start_end'].to(dev)
    n = int(sse[-1, 1])
    lab = data['obs_traj_g'][-1].to(dev)
    pos = data['obs_traj'][-1].to(dev)
    hbm = peaks.get('hbm_gbs', 6650.0)
    for name, mod in (('GCNModule', M.GCNModule()), ('GATEncoder', M.GATEncoder(None, 1, 0, 0.2))):
        mod = mod.to(dev)
        with torch.no_grad():
            for p in mod.parameters():
                if name == 'GCNModule' and p.dim() == 2 and tuple(p.shape) != (24, 32):
                    p.mul_(0.15)
        x = torch.randn(n, 40, device=dev)
        with torch.no_grad():
            f_ms = timed(lambda: mod(x, sse, pos, lab))
        xg = x.clone().requires_grad_(True)
        up = torch.randn(n, 24, device=dev)

        def fb():
            mod.zero_grad(set_to_none=True)
            xg.grad = None
            (mod(xg, sse, pos, lab) * up).sum().backward()
        fb_ms = timed(fb)
        fwd_b, bwd_b = 260 * n + 16 * (1 << 16), 420 * n
        print(json.dumps({'config': 'cfg3 op-level', 'op': name, 'scenes': 1 << 16, 'peds': n, 'fwd_ms': f_ms,
                          'fwd_bwd_ms': fb_ms, 'fwd_GBps_algorithmic': fwd_b / f_ms / 1e6,
                          'fwd_frac_hbm': fwd_b / f_ms / 1e6 / hbm,
                          'fwd_bwd_GBps_algorithmic': (fwd_b + bwd_b) / fb_ms / 1e6,
                          'fwd_bwd_frac_hbm': (fwd_b + bwd_b) / fb_ms / 1e6 / hbm}))


def cfg4():
    torch.manual_seed(0)
    gen = MD.TrajectoryGenerator(obs_len=8, pred_len=12, embedding_dim=16, encoder_h_dim=32, decoder_h_dim=32, mlp_dim=64,
                                 noise_dim=(8,), noise_mix_type='global', pooling_type='pool_net', pool_every_timestep=True,
                                 bottleneck_dim=8, batch_norm=False, n_heads=1, alpha=0.2).to(dev)
    for m in gen.modules():
        if isinstance(m, torch.nn.Linear):
            torch.nn.init.kaiming_normal_(m.weight)
    for precision in ('fp32', 'bf16', 'fp32-simt'):      # fp32 = the tcgen05 kernel with fp16 hi/lo operand splits
        gen.pool_net.precision = precision
        gen.decoder.pool_net.precision = precision
        for n in (64, 128, 256, 512, 1024):
            s = max(1, (8 << 20) // (n * n))
            rng = np.random.RandomState(n)
            sizes = [n] * s
            tot = n * s
            disp = rng.normal(0, 0.3, size=(8, tot, 2)).astype(np.float32)
            disp[0] = 0
            obs = rng.uniform(0, 15, size=(1, tot, 2)).astype(np.float32) + np.cumsum(disp, 0)
            lab = np.floor(rng.uniform(0, 1, tot) * max(1, n // 3)).astype(np.float32) + 1
            lab[rng.uniform(0, 1, tot) < 0.1] = 0
            t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
            o, orel = t(obs), t(disp)
            grp = t(np.broadcast_to(lab[None, :, None], (8, tot, 1)).copy())
            sse = sse_of(sizes).to(dev)
            z = torch.randn(s, 8, device=dev)
            with torch.no_grad():
                ms = timed(lambda: gen(o, orel, sse, grp, user_noise=z), reps=5, warm=2)
            pairs = n * n * s
            flops = 13 * pairs * bench.POOL_FLOPS_PER_PAIR
            print(json.dumps({'config': 'cfg4 dense crowd, pool_every_timestep=1, pred 12', 'precision': precision, 'N': n,
                              'scenes': s, 'peds': tot, 'pairs_per_pool_call': pairs, 'forward_ms': ms,
                              'traj_per_s_one_sample': tot / ms * 1e3,
                              'pool_as_written_TFLOPs_over_whole_forward': flops / ms / 1e9}))


def cfg4_ops():
    """Op-level GATEncoder / GCNModule / PoolHiddenNet at dense-crowd scene sizes (SURVEY 8d cfg 4): forward and
    forward+backward, algorithmic GB/s and FLOP/s (GAT: 190 N^2 + 8.4k N FLOP per scene, 260 B/ped fwd, 420 B/ped bwd)."""
    hbm = peaks.get('hbm_gbs', 6650.0)
    for n in (64, 128, 256, 512, 1024):
        s = max(1, 65536 // n)
        tot = n * s
        rng = np.random.RandomState(n)
        lab = np.floor(rng.uniform(0, 1, tot) * max(1, n // 3)).astype(np.float32) + 1
        lab[rng.uniform(0, 1, tot) < 0.1] = 0
        lab = torch.from_numpy(lab).view(-1, 1).to(dev)
        sse = sse_of([n] * s).to(dev)
        pos = torch.rand(tot, 2, device=dev) * 15
        for name, mod in (('GATEncoder', M.GATEncoder(None, 1, 0, 0.2)), ('GCNModule', M.GCNModule())):
            mod = mod.to(dev)
            with torch.no_grad():
                for p in mod.parameters():
                    if name == 'GCNModule' and p.dim() == 2 and tuple(p.shape) != (24, 32):
                        p.mul_(0.15)
            x = torch.randn(tot, 40, device=dev)
            with torch.no_grad():
                f_ms = timed(lambda: mod(x, sse, pos, lab), reps=5, warm=2)
            xg = x.clone().requires_grad_(True)
            up = torch.randn(tot, 24, device=dev)

            def fb():
                mod.zero_grad(set_to_none=True)
                xg.grad = None
                (mod(xg, sse, pos, lab) * up).sum().backward()
            fb_ms = timed(fb, reps=5, warm=2)
            flops = s * (190 * n * n + 8400 * n) if name == 'GATEncoder' else s * 8100 * n
            print(json.dumps({'config': 'cfg4 op-level', 'op': name, 'N': n, 'scenes': s, 'peds': tot, 'fwd_ms': f_ms,
                              'fwd_bwd_ms': fb_ms, 'fwd_GBps_algorithmic': 260 * tot / f_ms / 1e6,
                              'fwd_frac_hbm': 260 * tot / f_ms / 1e6 / hbm, 'fwd_GFLOPs_algorithmic': flops / f_ms / 1e6,
                              'fwd_bwd_GBps_algorithmic': 680 * tot / fb_ms / 1e6}))


def cfg2_small():
    """cfg 2 at the reference's own batch size: S = 64 zara1-shaped scenes (~240 peds), best-of-20 evaluation of one
    minibatch (scripts/evaluate_model.py:72-99).  Launch-latency regime: (a) the K forwards one after the other, as the
    reference issues them, (b) the K samples folded into ONE forward over 20 copies of the batch, (c) (b) replayed from
    a CUDA graph."""
    from group_gan_gcn_gat_b200.evaluate import evaluate_batch
    K = 20
    data = bench.synth_batch(64, 1238)
    gen = MD.TrajectoryGenerator(obs_len=8, pred_len=12, embedding_dim=16, encoder_h_dim=32, decoder_h_dim=32, mlp_dim=64,
                                 noise_dim=(8,), noise_mix_type='global', pooling_type='pool_net',
                                 pool_every_timestep=False, bottleneck_dim=8, batch_norm=False, n_heads=1, alpha=0.2)
    gen.load_state_dict(bench.load_weights(), strict=True)
    gen = gen.to(dev).train()
    d = {k: data[k].to(dev) for k in ('obs_traj', 'obs_traj_rel', 'obs_traj_g', 'seq_start_end', 'pred_traj_gt')}
    n = int(d['obs_traj'].shape[1])
    with torch.no_grad():
        loop_ms = timed(lambda: evaluate_batch(gen, d['obs_traj'], d['obs_traj_rel'], d['seq_start_end'], d['obs_traj_g'],
                                               d['pred_traj_gt'], K, fold_samples=False))
        fold_ms = timed(lambda: evaluate_batch(gen, d['obs_traj'], d['obs_traj_rel'], d['seq_start_end'], d['obs_traj_g'],
                                               d['pred_traj_gt'], K, fold_samples=True))
        # CUDA graph of the folded forward (static inputs; noise refreshed into a static buffer before each replay)
        s64 = d['seq_start_end'].shape[0]
        offs = (torch.arange(K, device=dev) * n).repeat_interleave(s64)
        sse_k = d['seq_start_end'].repeat(K, 1) + offs.unsqueeze(1)
        big = [d[k].repeat(1, K, 1).contiguous() for k in ('obs_traj', 'obs_traj_rel', 'obs_traj_g')]
        z = torch.randn(K * s64, 8, device=dev)
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(3):
                gen(big[0], big[1], sse_k, big[2], user_noise=z)
        torch.cuda.current_stream().wait_stream(side)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            out = gen(big[0], big[1], sse_k, big[2], user_noise=z)

        def replay():
            z.normal_()
            graph.replay()
        graph_ms = timed(replay)
    print(json.dumps({'config': 'cfg2 small batch: 64 zara1-shaped scenes, best-of-20', 'peds': n, 'k_samples': K,
                      'loop_of_20_forwards_ms': loop_ms, 'folded_one_forward_ms': fold_ms, 'folded_cuda_graph_ms': graph_ms,
                      'traj_per_s_loop': n * K / loop_ms * 1e3, 'traj_per_s_folded': n * K / fold_ms * 1e3,
                      'traj_per_s_graph': n * K / graph_ms * 1e3, 'graph_output_finite': bool(torch.isfinite(out).all())}))


if __name__ == '__main__':
    which = sys.argv[1:] or ['cfg2_small', 'cfg3', 'cfg4']
    if 'cfg2_small' in which:
        cfg2_small()
    if 'cfg3' in which:
        cfg3()
    if 'cfg4_ops' in which:
        cfg4_ops()
    if 'cfg4' in which:
        cfg4()
