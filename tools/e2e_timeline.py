"""Where one synchronised end-to-end evaluation step spends its time (bench.py `step_e2e`): host time stamps and CUDA events
at the seams of evaluate_batch on the default bench batch (SGAN-P, 65 536 scenes), plus the H2D bandwidth of the pinned
buffers.  python tools/e2e_timeline.py [--scenes N]"""
import argparse
import os
import statistics
import sys
import time

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--scenes', type=int, default=1 << 16)
    ap.add_argument('--config', default='sgan_p')
    args = ap.parse_args()
    dev = torch.device('cuda:0')
    torch.cuda.set_device(dev)
    from group_gan_gcn_gat_b200.evaluate import evaluate_batch
    from group_gan_gcn_gat_b200.schedule import get_schedule
    from group_gan_gcn_gat_b200.utils import stage_host_batch
    gen = bench.build_generator(args.config, dev)
    data = bench.synth_batch(args.scenes, 0, args.config)
    host = {k: data[k].pin_memory() for k in ('obs_traj', 'obs_traj_rel', 'obs_traj_g', 'seq_start_end', 'pred_traj_gt')}
    out_host = torch.empty(2).pin_memory()

    def ev():
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        return e

    # H2D bandwidth of one pinned tensor
    dst = torch.empty_like(host['obs_traj_rel'], device=dev)
    for _ in range(3):
        dst.copy_(host['obs_traj_rel'], non_blocking=True)
    torch.cuda.synchronize()
    a = ev()
    for _ in range(10):
        dst.copy_(host['obs_traj_rel'], non_blocking=True)
    b = ev()
    torch.cuda.synchronize()
    nbytes = dst.numel() * 4
    print('H2D pinned: %.1f MB in %.3f ms = %.1f GB/s' % (nbytes / 1e6, a.elapsed_time(b) / 10, nbytes * 10 / a.elapsed_time(b) / 1e6))

    with torch.no_grad():
        def step(record):
            marks = []

            def mark(name):
                if record:
                    marks.append((name, time.perf_counter(), ev()))
            mark('start')
            sse = host['seq_start_end'].clone()
            mark('sse clone')
            ade, fde = evaluate_batch(gen, host['obs_traj'], host['obs_traj_rel'], sse, host['obs_traj_g'],
                                      host['pred_traj_gt'], 20, fold_samples=False)
            mark('evaluate_batch returned (host) / all work queued')
            out_host.copy_(torch.stack([ade, fde]), non_blocking=True)
            mark('D2H queued')
            torch.cuda.synchronize()
            marks.append(('synchronised', time.perf_counter(), None))
            return marks

        for _ in range(5):
            step(False)
        rows = [step(True) for _ in range(10)]
        names = [m[0] for m in rows[0]]
        print('%-52s %10s %10s' % ('seam', 'host ms', 'device ms'))
        for i, n in enumerate(names):
            h = statistics.median((r[i][1] - r[0][1]) * 1e3 for r in rows)
            d = statistics.median(r[0][2].elapsed_time(r[i][2]) for r in rows) if rows[0][i][2] is not None else float('nan')
            print('%-52s %10.3f %10.3f' % (n, h, d))

        # host cost of the pieces in front of the first launch
        ts = []
        for _ in range(10):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            staged = stage_host_batch(dev, host['obs_traj'], host['obs_traj_rel'], host['obs_traj_g'], host['pred_traj_gt'])
            t1 = time.perf_counter()
            sched = get_schedule(host['seq_start_end'].clone(), dev)
            t2 = time.perf_counter()
            torch.cuda.synchronize()
            t3 = time.perf_counter()
            ts.append((t1 - t0, t2 - t1, t3 - t2))
            del staged, sched
        print('host: stage_host_batch %.3f ms, get_schedule %.3f ms, then %.3f ms until copies + schedule kernels are done'
              % tuple(statistics.median(x[i] for x in ts) * 1e3 for i in range(3)))

        # the device-resident step for comparison
        dev_in = {k: host[k].to(dev) for k in host}
        for _ in range(3):
            evaluate_batch(gen, dev_in['obs_traj'], dev_in['obs_traj_rel'], host['seq_start_end'].clone(), dev_in['obs_traj_g'],
                           dev_in['pred_traj_gt'], 20, fold_samples=False)
        torch.cuda.synchronize()
        ts = []
        for _ in range(10):
            t0 = time.perf_counter()
            a = ev()
            evaluate_batch(gen, dev_in['obs_traj'], dev_in['obs_traj_rel'], host['seq_start_end'].clone(), dev_in['obs_traj_g'],
                           dev_in['pred_traj_gt'], 20, fold_samples=False)
            b = ev()
            t1 = time.perf_counter()
            torch.cuda.synchronize()
            ts.append((a.elapsed_time(b), (t1 - t0) * 1e3, (time.perf_counter() - t0) * 1e3))
        print('device-resident inputs, same call: device %.3f ms, host issue %.3f ms, wall %.3f ms'
              % tuple(statistics.median(x[i] for x in ts) for i in range(3)))


if __name__ == '__main__':
    main()
