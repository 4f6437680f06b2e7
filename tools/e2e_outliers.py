"""Hunts the occasional slow e2e step of bench.py (one step of ~75 ms among ~11 ms ones): runs the same step_e2e for N
steps with host timestamps per phase and a gc callback, prints the outliers with the phase that stalled."""
import gc
import os
import statistics
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    n_steps = int(sys.argv[1]) if len(sys.argv) > 1 else 60
    freeze = (sys.argv[2] if len(sys.argv) > 2 else 'freeze') == 'freeze'
    from group_gan_gcn_gat_b200.evaluate import evaluate_batch
    torch.backends.cudnn.allow_tf32 = False
    dev = torch.device('cuda:0')
    data = bench.synth_batch(1 << 16, 1236, 'sgan_p')
    gen = bench.build_generator('sgan_p', dev)
    host = {k: data[k].pin_memory() for k in ('obs_traj', 'obs_traj_rel', 'obs_traj_g', 'seq_start_end', 'pred_traj_gt')}
    out_host = torch.empty(2).pin_memory()
    gc_log = []
    t_gc = [0.0]

    def cb(phase, info):
        if phase == 'start':
            t_gc[0] = time.perf_counter()
        else:
            gc_log.append((time.perf_counter(), (time.perf_counter() - t_gc[0]) * 1e3, info['generation'], info['collected']))
    gc.callbacks.append(cb)
    if freeze:
        gc.collect()
        gc.freeze()
    rows = []
    # timestamps around every generator forward and around every library call made through ctypes
    marks = []
    orig_forward = gen.forward

    def timed_forward(*a, **k):
        t0 = time.perf_counter()
        r = orig_forward(*a, **k)
        marks.append(('forward', t0, time.perf_counter()))
        return r
    gen.forward = timed_forward
    import torch.nn.functional as F  # noqa: F401
    slow_calls = []
    prof_on = os.environ.get('E2E_TRACE') == '1'
    if prof_on:
        import sys as _sys

        def tracer(frame, event, arg):
            if event == 'c_call' or event == 'call':
                frame.f_locals  # noqa: B018
            return None
    dev_allocs = []
    with torch.no_grad():
        for i in range(n_steps + 5):
            ms = torch.cuda.memory_stats(dev)
            dev_allocs.append((ms.get('num_device_alloc', 0), ms.get('num_device_free', 0), ms.get('reserved_bytes.all.current', 0) >> 20))
            t = [time.perf_counter()]
            d = {k: host[k].to(dev, non_blocking=True) for k in ('obs_traj', 'obs_traj_rel', 'obs_traj_g', 'pred_traj_gt')}
            t.append(time.perf_counter())
            sse = host['seq_start_end'].clone()
            t.append(time.perf_counter())
            ade, fde = evaluate_batch(gen, d['obs_traj'], d['obs_traj_rel'], sse, d['obs_traj_g'], d['pred_traj_gt'], 20,
                                      fold_samples=False)
            t.append(time.perf_counter())
            out_host.copy_(torch.stack([ade, fde]), non_blocking=True)
            torch.cuda.synchronize()
            t.append(time.perf_counter())
            if i >= 5:
                rows.append(t)
                fw = [(b - a) * 1e3 for (_n, a, b) in marks if t[0] <= a <= t[-1]]
                gaps = [(b[1] - a[2]) * 1e3 for a, b in zip(marks, marks[1:]) if t[0] <= a[1] and b[2] <= t[-1]]
                if (t[-1] - t[0]) > 0.016:
                    slow_calls.append((i, [round(v, 1) for v in fw], [round(v, 1) for v in gaps]))
            marks.clear()
    tot = [(r[-1] - r[0]) * 1e3 for r in rows]
    med = statistics.median(tot)
    print('steps %d median %.2f ms mean %.2f max %.2f (gc %s)' % (len(tot), med, statistics.mean(tot), max(tot),
                                                                   'frozen' if freeze else 'default'))
    names = ['h2d issue', 'sse clone', 'evaluate_batch issue', 'd2h + drain']
    for r, tt in zip(rows, tot):
        if tt > 1.5 * med:
            gcs = [(round(ms, 1), gen_, n) for (ts, ms, gen_, n) in gc_log if r[0] <= ts <= r[-1] + 1e-3]
            print('  outlier %.1f ms: ' % tt + ', '.join('%s %.1f' % (n, (b - a) * 1e3) for n, a, b in zip(names, r, r[1:])),
                  ' gc inside:', gcs)
    for i, fw, gaps in slow_calls[:6]:
        print('  step %d forwards (ms): %s\n      gaps between forwards (ms): %s' % (i, fw, gaps))
    print('allocator (cudaMalloc calls, cudaFree calls, reserved MiB) at steps 0, 5, 10, 20, 40, last:',
          [dev_allocs[k] for k in (0, 5, 10, 20, 40, len(dev_allocs) - 1) if k < len(dev_allocs)])
    print('gc events total:', len(gc_log), 'slowest:', sorted(gc_log, key=lambda g: -g[1])[:3])


if __name__ == '__main__':
    main()
