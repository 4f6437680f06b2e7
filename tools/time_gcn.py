"""Time the GCNModule forward alone on the bench workload (CUDA events, L2 flushed between launches).
usage: [SGX_LIB=variant.so] python tools/time_gat.py [scenes]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    scenes = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 16
    from group_gan_gcn_gat_b200 import modules as M
    dev = torch.device('cuda:0')
    data = bench.synth_batch(scenes, 1236)
    enc = M.GCNModule()
    torch.manual_seed(0)
    with torch.no_grad():
        for p in enc.parameters():
            if p.dim() == 2 and tuple(p.shape) != (24, 32):
                p.mul_(0.15)
    enc = enc.to(dev)
    sse = data['seq_start_end'].to(dev)
    batch = int(data['obs_traj'].shape[1])
    g = torch.Generator(device='cpu').manual_seed(1)
    h = torch.randn(batch, 40, generator=g).to(dev)
    end_pos = data['obs_traj'][-1].to(dev)
    end_group = data['obs_traj_g'][-1].to(dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(20)]
    with torch.no_grad():
        for _ in range(3):
            out = enc(h, sse, end_pos, end_group)
        for a, b in ev:
            flush.zero_()
            a.record()
            out = enc(h, sse, end_pos, end_group)
            b.record()
        torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in ev)
    print('lib=%s peds=%d gcn_module fwd (groups + kernels): median %.1f us  min %.1f us  checksum %.6f'
          % (os.path.basename(os.environ.get('SGX_LIB', 'default')), batch, ts[len(ts) // 2] * 1e3, ts[0] * 1e3,
             float(out.double().sum())))


if __name__ == '__main__':
    main()
