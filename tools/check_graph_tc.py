"""tcgen05 GATEncoder / GCNModule forwards against the mma.sync kernels on the bench workload: max relative difference
and time of both (CUDA events, L2 flushed).  usage: python tools/check_graph_tc.py [scenes]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def timed(fn, flush, reps=20):
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
    for _ in range(3):
        fn()
    for a, b in ev:
        flush.zero_()
        a.record()
        fn()
        b.record()
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in ev)
    return ts[len(ts) // 2] * 1e3, ts[0] * 1e3


def main():
    scenes = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 16
    from group_gan_gcn_gat_b200 import modules as M, _lib
    dev = torch.device('cuda:0')
    data = bench.synth_batch(scenes, 1236)
    sd = {k[len('gatencoder.'):]: v for k, v in bench.load_weights().items() if k.startswith('gatencoder.')}
    enc = M.GATEncoder(n_units=[40, 72, 16], n_heads=1, dropout=0, alpha=0.2)
    enc.load_state_dict(sd, strict=True)
    enc = enc.to(dev)
    gcn = M.GCNModule().to(dev)
    with torch.no_grad():
        for prm in gcn.parameters():
            if prm.dim() == 2 and tuple(prm.shape) != (24, 32):
                prm.mul_(0.15)
    sse = data['seq_start_end'].to(dev)
    batch = int(data['obs_traj'].shape[1])
    g = torch.Generator(device='cpu').manual_seed(1)
    h = torch.randn(batch, 40, generator=g).to(dev)
    end_pos = data['obs_traj'][-1].to(dev)
    end_group = data['obs_traj_g'][-1].to(dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    with torch.no_grad():
        for name, mod in (('gat_encoder', enc), ('gcn_module', gcn)):
            if mod is None:
                continue
            for scale in (1.0, 1e-3, 1e3):
                hs = h * scale
                fn = lambda: mod(hs, sse, end_pos, end_group)
                _lib.set_option('graph_tc', 0)
                ref = fn().clone()
                _lib.set_option('graph_tc', 1)
                out = fn().clone()
                torch.cuda.synchronize()
                err = float((out - ref).abs().max() / ref.abs().max())
                print('%s scale %g: max |tc - mma| / max |mma| = %.3e  (max |ref| %.4g, nan %d)'
                      % (name, scale, err, float(ref.abs().max()), int(torch.isnan(out).sum())), flush=True)
            fn = lambda: mod(h, sse, end_pos, end_group)
            _lib.set_option('graph_tc', 0)
            t0 = timed(fn, flush)
            _lib.set_option('graph_tc', 1)
            t1 = timed(fn, flush)
            print('%s peds=%d fwd (groups + kernel): mma.sync median %.1f us min %.1f | tcgen05 median %.1f us min %.1f'
                  % (name, batch, t0[0], t0[1], t1[0], t1[1]), flush=True)


if __name__ == '__main__':
    main()
