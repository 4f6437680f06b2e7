#!/usr/bin/env python
"""bench.py -- predicted trajectories/sec of the sgan generator forward (K = 20 samples) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl sgx|reference] [--config sgan_p|sgan_gat]
                    [--scenes S] [--precision fp32|fp32-simt|tc32|bf16]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workloads (BASELINE.json `metric` is quoted on "SGAN-P fwd, K=20" = configs[0]; SURVEY.md 8d cfg 1 / cfg 2):
  sgan_p   (default) SGAN-P generator: encoder LSTM -> PoolHiddenNet -> mlp_decoder_context -> noise -> decoder LSTM
           (the upstream wiring, sgan/models.py:796-804,898), weights of models/sgan-p-models/eth_8_model.pt (frozen in
           tests/golden/generator_p_eth.npz), obs 8 / pred 8, scene sizes from the ETH-test histogram tiled to S = 2^16
           scenes per GPU (SURVEY 8d cfg 1 scale-up).
  sgan_gat SGAN-GAT generator (PoolHiddenNet + GATEncoder), zara1-test histogram, obs 8 / pred 12,
           models/sgan-gat-models/zara1_12_model.pt (tests/golden/generator_gat_zara1.npz).
One *step* = the K=20 best-of-K generator forwards over one batch of S scenes per GPU (the loop of
scripts/evaluate_model.py:85-90); every forward is complete (nothing is hoisted out of the K loop).
trajectories/step = peds * 20.  Pooling runs in the default 'fp32' mode -- the mode whose ADE/FDE <= 1e-4 tests are
green (tests/test_gpu_parity.py::test_generator_matches_reference): the tcgen05 kernel with fp16 hi/lo operand splits.

Printed JSON line (rank 0): value = whole-job trajectories/sec with inputs resident in HBM; e2e = the same
through the module API from pinned HOST buffers (schedule build + H2D + 20 forwards + D2H inside the timed
region); roofline = the PoolHiddenNet pair kernel (CUDA events recorded by the library around that kernel);
cpu_baseline = the CPU oracle port of the reference timed on this box's host cores on a bounded sample.
--impl reference times that CPU port alone (the reference itself is Python and cannot travel to the GPU box), on the
same config: each step is a bounded sample (--ref-scenes scenes of the same histogram) of the workload.
"""
import argparse
import gc
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ZARA1_HIST = {2: 212, 3: 136, 4: 109, 5: 55, 6: 32, 7: 10, 8: 20, 9: 8, 10: 12, 11: 4, 12: 1, 13: 2, 14: 1}
ETH_HIST = {2: 89, 3: 53, 4: 22, 5: 19, 6: 7, 7: 1, 8: 1, 12: 2, 13: 1}          # SURVEY 8d cfg 1 / A.3 (eth test, pred 8)
K_SAMPLES = 20
OBS_LEN, PRED_LEN = 8, 12            # sgan_gat defaults (kept as module constants for tools/ and tests/)
POOL_FLOPS_PER_PAIR = 4 * 16 + 2 * (16 + 32) * 512 + 2 * 512 * 8     # 57 408, as written
# tensor-pipe FLOPs actually issued per ordered pair (padded N, operand-split K): DESIGN.md 4.1 / 4.1b
POOL_EXECUTED_FLOPS = {'bf16': 2 * 512 * 48 + 2 * 16 * 512, 'tc32': 2 * 512 * 112 + 2 * 16 * 1024,
                       'fp32-simt': 512 * (2 * 2 + 1 + 2 * 8)}

CONFIGS = {
    'sgan_p': dict(hist=ETH_HIST, pred_len=8, wiring='mlp', golden='generator_p_eth',
                   what='SGAN-P generator fwd (encoder LSTM, PoolHiddenNet, mlp_decoder_context, decoder LSTM), '
                        'ETH-test scene-size histogram tiled, obs 8 / pred 8, K=20 forwards per step',
                   weights='models/sgan-p-models/eth_8_model.pt (tests/golden/generator_p_eth.npz)'),
    'sgan_gat': dict(hist=ZARA1_HIST, pred_len=12, wiring='gat', golden='generator_gat_zara1',
                     what='SGAN-GAT generator fwd (PoolHiddenNet+GATEncoder), zara1-shaped synthetic scenes, '
                          'obs 8 / pred 12, K=20 forwards per step',
                     weights='models/sgan-gat-models/zara1_12_model.pt (tests/golden/generator_gat_zara1.npz)'),
}


def synth_batch(n_scenes, seed, config='sgan_gat'):
    """positions U[0,15]^2, per-step displacement N(0,0.3^2), labels 10% zero else U{1..max(1,N//3)};
    scene sizes drawn from the config's histogram (ETH test for sgan_p, zara1 test for sgan_gat)."""
    cfg = CONFIGS[config]
    pred_len = cfg['pred_len']
    rng = np.random.RandomState(seed)
    sizes_pool = np.array(list(cfg['hist'].keys()))
    probs = np.array(list(cfg['hist'].values()), dtype=np.float64)
    sizes = rng.choice(sizes_pool, size=n_scenes, p=probs / probs.sum())
    starts = np.concatenate([[0], np.cumsum(sizes)])
    batch = int(starts[-1])
    sse = np.stack([starts[:-1], starts[1:]], axis=1).astype(np.int64)
    disp = rng.normal(0, 0.3, size=(OBS_LEN, batch, 2)).astype(np.float32)
    disp[0] = 0
    p0 = rng.uniform(0, 15, size=(1, batch, 2)).astype(np.float32)
    obs = p0 + np.cumsum(disp, axis=0)
    hi = np.maximum(1, np.repeat(sizes, sizes) // 3)
    lab = np.floor(rng.uniform(0, 1, size=batch) * hi).astype(np.float32) + 1
    lab[rng.uniform(0, 1, size=batch) < 0.10] = 0
    grp = np.broadcast_to(lab[None, :, None], (OBS_LEN, batch, 1)).copy()
    fut = obs[-1:] + np.cumsum(rng.normal(0, 0.3, size=(pred_len, batch, 2)).astype(np.float32), axis=0)
    return dict(pred_traj_gt=torch.from_numpy(fut.astype(np.float32)),
                obs_traj=torch.from_numpy(obs.astype(np.float32)), obs_traj_rel=torch.from_numpy(disp),
                obs_traj_g=torch.from_numpy(grp), seq_start_end=torch.from_numpy(sse), sizes=sizes)


def load_weights(config='sgan_gat'):
    z = np.load(os.path.join(ROOT, 'tests', 'golden', CONFIGS[config]['golden'] + '.npz'))
    return {k[3:]: torch.from_numpy(z[k]) for k in z.files if k.startswith('sd.')}


def build_generator(config, dev):
    """The generator of the config with the shipped checkpoint's weights, in train mode (scripts/evaluate_model.py:54)."""
    from group_gan_gcn_gat_b200 import models as MD
    cfg = CONFIGS[config]
    gen = MD.TrajectoryGenerator(obs_len=OBS_LEN, pred_len=cfg['pred_len'], embedding_dim=16, encoder_h_dim=32,
                                 decoder_h_dim=32, mlp_dim=64, noise_dim=(8,), noise_mix_type='global',
                                 pooling_type='pool_net', pool_every_timestep=False, bottleneck_dim=8, batch_norm=False,
                                 n_heads=1, alpha=0.2, context_type=cfg['wiring'])
    sd = load_weights(config)
    if cfg['wiring'] == 'gat':
        gen.load_state_dict(sd, strict=True)
    else:   # older checkpoint generation: the reference class's passenger gcn_module is not part of this wiring
        missing, unexpected = gen.load_state_dict(sd, strict=False)
        assert not missing and all(k.startswith('gcn_module') for k in unexpected), (missing, unexpected)
    return gen.to(dev).train()


def bind_to_gpu_numa(local_rank):
    """Pin this rank's host threads to the NUMA node of its GPU (H2D staging and the schedule build are host work)."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local_rank)
        bus = pynvml.nvmlDeviceGetPciInfo(h).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        node = int(open('/sys/bus/pci/devices/%s/numa_node' % bus[-12:].lower()).read())
        if node < 0:
            return None
        cpus = []
        for part in open('/sys/devices/system/node/node%d/cpulist' % node).read().strip().split(','):
            a, _, b = part.partition('-')
            cpus += list(range(int(a), int(b or a) + 1))
        allowed = sorted(set(cpus) & os.sched_getaffinity(0))
        if allowed:
            os.sched_setaffinity(0, allowed)
            return node
    except Exception:
        return None
    return None


class ClockSampler:
    QUERY = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,'
             'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
             'clocks_event_reasons.sw_power_cap')

    def __init__(self, gpu_index):
        self.rows, self.proc, self.gpu = [], None, gpu_index

    def start(self):
        """NVML polled every 8 ms from a thread (the timed region of a default run is ~200 ms, about one
        `nvidia-smi -lms` period); the recipe's nvidia-smi query line is the fallback."""
        self.stop_flag = threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(self.gpu)
            reasons_fn = getattr(pynvml, 'nvmlDeviceGetCurrentClocksEventReasons', None) or \
                pynvml.nvmlDeviceGetCurrentClocksThrottleReasons
            bits = [('hw_slowdown', 0x8), ('hw_thermal_slowdown', 0x40), ('sw_thermal_slowdown', 0x20), ('sw_power_cap', 0x4)]
            max_sm = pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)

            def poll():
                while not self.stop_flag.is_set():
                    try:
                        sm = pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)
                        r = reasons_fn(h)
                        pw = pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0
                    except pynvml.NVMLError:
                        break
                    self.rows.append([time.perf_counter(), str(self.gpu), str(sm), str(max_sm), '%.1f' % pw, hex(r)] +
                                     ['Active' if r & b else 'Not Active' for _n, b in bits])
                    self.stop_flag.wait(0.008)

            self.proc = 'nvml'
            self.source = 'nvml, 8 ms period'
            threading.Thread(target=poll, daemon=True).start()
            return
        except Exception:
            pass
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.gpu), '--query-gpu=' + self.QUERY,
                                          '--format=csv,noheader,nounits', '-lms', '100'], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.source = 'nvidia-smi -lms 100'
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([time.perf_counter()] + [c.strip() for c in line.split(',')])

    def window(self, t0, t1):
        """keep the samples taken inside [t0, t1] (the timed region); fall back to everything if there are < 3"""
        inside = [r[1:] for r in self.rows if t0 <= r[0] <= t1]
        self.note = 'samples inside the timed region' if len(inside) >= 3 else \
            'timed region shorter than 3 sampling periods: includes warm-up samples'
        self.rows = inside if len(inside) >= 3 else [r[1:] for r in self.rows]

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        self.stop_flag.set()
        if self.proc != 'nvml':
            self.proc.terminate()
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace('.', '').isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace('.', '').isdigit()]
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        reasons = sorted({n for r in self.rows if len(r) >= 9 for n, v in zip(names, r[5:9]) if v.lower() == 'active'})
        return {'sm_mhz': statistics.median(sm) if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'reasons': reasons, 'samples': len(sm), 'note': getattr(self, 'note', ''),
                'source': getattr(self, 'source', '')}


def cpu_port_traj_per_sec(n_scenes, k_samples, seed, reps=1, config='sgan_gat'):
    """The CPU oracle port of the reference generator (per-scene python loop, N^2 materialisation), all host threads."""
    from oracle import sgan_oracle as O
    torch.set_num_threads(os.cpu_count() or 1)
    cfgd = CONFIGS[config]
    data = synth_batch(n_scenes, seed, config)
    sd = load_weights(config)
    cfg = dict(pred_len=cfgd['pred_len'], wiring=cfgd['wiring'], pooling=True, pool_every_timestep=False, alpha=0.2, n_heads=1)
    gen = torch.Generator().manual_seed(seed)
    best = None
    with torch.no_grad():
        for _ in range(reps):
            t0 = time.perf_counter()
            ade, fde = [], []
            for _k in range(k_samples):              # the loop of scripts/evaluate_model.py:85-95
                z = torch.randn(n_scenes, 8, generator=gen)
                rel = O.generator_forward(data['obs_traj'], data['obs_traj_rel'], data['seq_start_end'],
                                          data['obs_traj_g'], sd, cfg, z)
                pred = O.relative_to_abs(rel, data['obs_traj'][-1])
                ade.append(O.displacement_error_raw(pred, data['pred_traj_gt']))
                fde.append(O.final_displacement_error_raw(pred[-1], data['pred_traj_gt'][-1]))
            O.best_of_k(ade, data['seq_start_end'])
            O.best_of_k(fde, data['seq_start_end'])
            dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
    peds = int(data['seq_start_end'][-1, 1])
    return peds * k_samples / best, best, peds


def workload_config(config, n_scenes):
    """Identical for both arms: the reference arm runs bounded samples of THIS workload (see its cpu_baseline.sample)."""
    cfg = CONFIGS[config]
    return {'workload': cfg['what'], 'name': config, 'scenes_per_gpu': n_scenes, 'k_samples': K_SAMPLES,
            'obs_len': OBS_LEN, 'pred_len': cfg['pred_len'], 'pool_precision': 'fp32', 'weights': cfg['weights'],
            'l2': 'flushed between timed steps (256 MiB write)', 'parallelism': 'scenes sharded by LPT on N^2'}


def run_reference(args, rank):
    if rank != 0:
        return
    n_scenes = args.ref_scenes
    for _ in range(args.warmup):
        cpu_port_traj_per_sec(16, 2, 1, config=args.config)
    t0 = time.perf_counter()
    vals = []
    for s in range(args.steps):
        v, dt, peds = cpu_port_traj_per_sec(n_scenes, K_SAMPLES, 1234 + 2 + s, config=args.config)
        vals.append((v, dt, peds))
    total_traj = sum(p * K_SAMPLES for _, _, p in vals)
    total_t = sum(dt for _, dt, _ in vals)
    value = total_traj / total_t
    line = {'impl': 'reference', 'metric': 'predicted_trajectories_per_sec', 'value': value, 'unit': 'traj/s',
            'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': 1e3 * total_t / args.steps,
            'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'fp32', 'data': 'synthetic',
            'config': workload_config(args.config, args.scenes),
            'cpu_baseline': {'value': value, 'unit': 'traj/s', 'cores': os.cpu_count(), 'kind': 'port',
                             'sample': 'each step = %d scenes of the config\'s histogram x K=%d complete generator forwards '
                                       '(oracle port of sgan/models.py, per-scene loop, fp32, torch threads = all cores); '
                                       'the reference has no batching across scenes, so traj/s does not depend on the '
                                       'sample size' % (n_scenes, K_SAMPLES)},
            'e2e': {'value': value, 'unit': 'traj/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
            'wall_s': time.perf_counter() - t0}
    print(json.dumps(line))


def nccl_evidence(path):
    """what NCCL itself logged about the communicator of this rank: rank count, NVLS, transport"""
    if not path:
        return {'log': None, 'note': 'NCCL_DEBUG=%s was set by the caller; its log is wherever the caller sent it (%s)'
                             % (os.environ.get('NCCL_DEBUG'), os.environ.get('NCCL_DEBUG_FILE', 'stdout'))}
    try:
        text = open(path, errors='replace').read()
    except OSError:
        return {'log': os.path.relpath(path, ROOT), 'note': 'log not found'}
    import re
    nranks = sorted({int(m) for m in re.findall(r'nranks (\d+)', text)})
    return {'log': os.path.relpath(path, ROOT), 'nranks': nranks[-1] if nranks else None,
            'nvls': ('NVLS' in text), 'p2p_nvlink': ('via P2P' in text or 'NVL' in text),
            'version': (re.findall(r'NCCL version ([0-9.+a-z]+)', text) or [None])[0]}


def pool_kernel_name(code):
    return {0: 'pool_pair_kernel', 1: 'pool_tc_kernel', 2: 'pool_tc32_kernel'}[code]


def measured_traffic(config, kernel, scenes):
    """dram bytes per launch of the dominant kernel from the committed ncu summary (profiles/pool_traffic.json, written
    by tools/ncu_traffic.py from an `ncu --set full` capture of this command); None when no capture matches."""
    try:
        table = json.load(open(os.path.join(ROOT, 'profiles', 'pool_traffic.json')))
    except (OSError, ValueError):
        return None, None
    e = table.get('%s/%s/%d' % (config, kernel, scenes))
    return (e['dram_bytes_per_launch'], e['source']) if e else (None, None)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='sgx', choices=['sgx', 'reference'])
    ap.add_argument('--config', default='sgan_p', choices=sorted(CONFIGS))
    ap.add_argument('--scenes', type=int, default=1 << 16, help='scenes per GPU per step')
    ap.add_argument('--precision', default=os.environ.get('SGX_POOL_PRECISION', 'fp32'),
                    help="pooling precision: fp32 (default, contract mode) | fp32-simt | tc32 | bf16 (outside the ADE contract)")
    ap.add_argument('--mode', default='infer', choices=['infer', 'train'],
                    help="infer (default): the headline best-of-K generator forwards; train: SURVEY cfg 5, the adversarial "
                         "D + G iteration data-parallel over scenes as its own JSON line")
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-train', action='store_true', help='skip the cfg-5 training-step sub-measurement')
    ap.add_argument('--ref-scenes', type=int, default=256,
                    help='scenes per step of the --impl reference arm (a bounded sample of the same workload)')
    args = ap.parse_args()
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    if args.impl == 'reference':
        run_reference(args, rank)
        return
    args.warmup = max(args.warmup, 3)
    cfg = CONFIGS[args.config]
    pred_len = cfg['pred_len']

    import torch.distributed as dist
    assert torch.cuda.is_available(), 'bench.py needs a CUDA device (the sgx ops have no CPU fallback)'
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    numa = bind_to_gpu_numa(local_rank)
    nccl_log = None
    if world > 1:
        # NCCL's own account of the communicator (ranks, NVLS) goes to a file per rank -- never to stdout, which carries
        # the JSON line; a caller's own NCCL_DEBUG settings are left alone
        if os.environ.get('NCCL_DEBUG', '').upper() not in ('INFO', 'TRACE'):      # unset, or VERSION / WARN from the image
            os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
            nccl_log = os.path.join(ROOT, 'gpurun_out', 'nccl_n%d_pid%d.log' % (world, os.getpid()))
            os.environ['NCCL_DEBUG'] = 'INFO'
            os.environ['NCCL_DEBUG_SUBSYS'] = 'INIT,GRAPH'
            os.environ['NCCL_DEBUG_FILE'] = nccl_log
        dist.init_process_group('nccl', device_id=dev)
    torch.backends.cudnn.allow_tf32 = False

    from group_gan_gcn_gat_b200 import _lib, modules as M
    from group_gan_gcn_gat_b200.schedule import SceneSchedule
    L = _lib.lib()
    if args.mode == 'train':
        run_train_mode(args, dev, rank, world, local_rank, nccl_log, L)
        if world > 1:
            dist.destroy_process_group()
        return
    precision = args.precision
    pool_code = M.resolve_pool_precision(precision, 16, 32, 8)
    kernel = pool_kernel_name(pool_code)
    pool_mode = {0: 'fp32-simt', 1: 'bf16', 2: 'tc32'}[pool_code]

    # ---- the global scene set, sharded by LPT on N^2 (weak scaling: S scenes per GPU) ----
    data = synth_batch(args.scenes * world, 1234 + 2, args.config)
    if world > 1:
        full = SceneSchedule(data['seq_start_end'], 'cpu')
        rank_of, _ = full.partition(world)
        mine = np.nonzero(rank_of == rank)[0]
        sse = data['seq_start_end'].numpy()
        idx = np.concatenate([np.arange(sse[s, 0], sse[s, 1]) for s in mine])
        sizes = data['sizes'][mine]
        st = np.concatenate([[0], np.cumsum(sizes)])
        data = dict(obs_traj=data['obs_traj'][:, idx], obs_traj_rel=data['obs_traj_rel'][:, idx],
                    obs_traj_g=data['obs_traj_g'][:, idx], pred_traj_gt=data['pred_traj_gt'][:, idx],
                    seq_start_end=torch.from_numpy(np.stack([st[:-1], st[1:]], 1).astype(np.int64)), sizes=sizes)
    n_scenes = data['seq_start_end'].shape[0]
    peds = int(data['seq_start_end'][-1, 1])
    n_pairs = int((data['sizes'].astype(np.int64) ** 2).sum())

    gen = build_generator(args.config, dev)
    gen.pool_net.precision = precision

    host = {k: data[k].pin_memory() for k in ('obs_traj', 'obs_traj_rel', 'obs_traj_g', 'seq_start_end', 'pred_traj_gt')}
    dev_in = {k: v.to(dev) for k, v in host.items()}
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    noise_gen = torch.Generator(device=dev).manual_seed(7)
    out_host = torch.empty(2).pin_memory()
    from group_gan_gcn_gat_b200.evaluate import evaluate_batch

    def step_resident():
        outs = None
        for _k in range(K_SAMPLES):
            z = torch.randn(n_scenes, 8, device=dev, generator=noise_gen)
            outs = gen(dev_in['obs_traj'], dev_in['obs_traj_rel'], dev_in['seq_start_end'], dev_in['obs_traj_g'],
                       user_noise=z)
        return outs

    def step_e2e():
        """The body of scripts/evaluate_model.py:72-99 for one minibatch, from HOST buffers: H2D of the batch, schedule
        built from the host seq_start_end, K complete generator forwards, best-of-K ADE/FDE reduced on the device,
        D2H of the two sums."""
        sse = host['seq_start_end'].clone()              # a fresh batch object every step: the schedule is rebuilt
        # pinned HOST tensors straight into the public call: evaluate_batch stages them (copy stream, obs_traj_rel first,
        # one event per tensor, the compute stream waits for each where the forward first reads it)
        ade, fde = evaluate_batch(gen, host['obs_traj'], host['obs_traj_rel'], sse, host['obs_traj_g'],
                                  host['pred_traj_gt'], K_SAMPLES, fold_samples=False)
        out_host.copy_(torch.stack([ade, fde]), non_blocking=True)
        torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # Long-lived objects (torch, the modules, the batch) go to the permanent generation: a full collection of the
    # interpreter's ~10^6 import-time objects in the middle of a step costs 20-40 ms of launch-issue time.
    gc.collect()
    gc.freeze()
    with torch.no_grad():
        sampler = ClockSampler(local_rank)
        if rank == 0:
            sampler.start()
        for _ in range(args.warmup):
            step_resident()
        # ---- timed: resident inputs, CUDA events per step, L2 flushed between steps ----
        barrier()
        launches0 = L.sgx_launch_count()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
        wall0 = time.perf_counter()
        for a, b in ev:
            flush.fill_(1)
            a.record()
            step_resident()
            b.record()
        barrier()
        wall = time.perf_counter() - wall0
        if rank == 0:
            sampler.window(wall0, wall0 + wall)
        launches = L.sgx_launch_count() - launches0
        step_ms = [a.elapsed_time(b) for a, b in ev]
        total_ms = sum(step_ms)
        clocks = sampler.stop() if rank == 0 else None

        # ---- e2e: host buffers, H2D + schedule + 20 forwards + D2H inside the timed region ----
        for _ in range(args.warmup + 2):                 # the first e2e steps grow the caching allocator and the pinned
            step_e2e()                                   # staging buffers of the schedule; keep that out of the timing
        barrier()
        t0 = time.perf_counter()
        e2e_steps = []
        for _ in range(args.steps):
            t1 = time.perf_counter()
            step_e2e()
            e2e_steps.append((time.perf_counter() - t1) * 1e3)
        barrier()
        e2e_s = time.perf_counter() - t0

        # ---- the same K steps as a pipeline (what evaluate() over data.DeviceLoader does): batch k+1 is copied on a side
        # stream while batch k computes, results go back with an async D2H per step, ONE synchronisation at the end --
        # the reference's evaluate() does not synchronise per minibatch either (scripts/evaluate_model.py:72-99).
        side = torch.cuda.Stream(dev)
        main_stream = torch.cuda.current_stream(dev)
        out_pipe = torch.empty(args.steps + args.warmup, 2).pin_memory()
        keys = ('obs_traj_rel', 'obs_traj', 'obs_traj_g', 'pred_traj_gt')
        slots = [{k: torch.empty_like(dev_in[k]) for k in keys} for _ in range(2)]    # double buffer, no allocation per step

        def stage(i):
            side.wait_stream(main_stream)            # slot i % 2 was last read by step i - 2, already enqueued on main
            with torch.cuda.stream(side):
                for k in keys:
                    slots[i % 2][k].copy_(host[k], non_blocking=True)
                done = torch.cuda.Event()
                done.record(side)
            return slots[i % 2], done

        pipe_marks = []

        def run_pipelined(n, offset):
            nxt = stage(0)
            first = torch.cuda.Event(enable_timing=True)
            first.record(main_stream)
            pipe_marks[:] = [(time.perf_counter(), first)]
            for i in range(n):
                cur, done = nxt
                main_stream.wait_event(done)
                if i + 1 < n:
                    nxt = stage(i + 1)
                sse = host['seq_start_end'].clone()
                ade, fde = evaluate_batch(gen, cur['obs_traj'], cur['obs_traj_rel'], sse, cur['obs_traj_g'],
                                          cur['pred_traj_gt'], K_SAMPLES, fold_samples=False)
                out_pipe[offset + i].copy_(torch.stack([ade, fde]), non_blocking=True)
                mark = torch.cuda.Event(enable_timing=True)
                mark.record(main_stream)
                pipe_marks.append((time.perf_counter(), mark))

        run_pipelined(args.warmup, 0)
        barrier()
        t0 = time.perf_counter()
        run_pipelined(args.steps, args.warmup)
        barrier()
        pipe_s = time.perf_counter() - t0
        pipe_dev_ms = [a[1].elapsed_time(b[1]) for a, b in zip(pipe_marks, pipe_marks[1:])]
        pipe_host_ms = [(b[0] - a[0]) * 1e3 for a, b in zip(pipe_marks, pipe_marks[1:])]

        # ---- roofline of the dominant pooling kernel: events recorded by the library around that launch ----
        import ctypes
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); e1.record(); torch.cuda.synchronize()        # create the underlying cudaEvents
        L.sgx_profile_events(ctypes.c_void_p(e0.cuda_event), ctypes.c_void_p(e1.cuda_event))
        h_enc = gen.encoder(dev_in['obs_traj_rel'])
        kernel_ms = []
        for _ in range(3 + 10):
            flush.fill_(1)
            gen.pool_net(h_enc, dev_in['seq_start_end'], dev_in['obs_traj'][-1])
            torch.cuda.synchronize()
            kernel_ms.append(e0.elapsed_time(e1))
        L.sgx_profile_events(None, None)
        kernel_ms = kernel_ms[3:]

        # ---- the other hot-path ops, CUDA events around the module call (op = its handful of launches) ----
        def time_call(fn, reps=10):
            ts = []
            for i in range(reps + 3):
                flush.fill_(i)
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                fn()
                b.record()
                torch.cuda.synchronize()
                ts.append(a.elapsed_time(b))
            return statistics.mean(ts[3:])
        end_pos = dev_in['obs_traj'][-1]
        pool_ms = time_call(lambda: gen.pool_net(h_enc, dev_in['seq_start_end'], end_pos))
        pool_h = gen.pool_net(h_enc, dev_in['seq_start_end'], end_pos)
        ctx_in = torch.cat([h_enc.view(-1, 32), pool_h], dim=1)
        end_grp = dev_in['obs_traj_g'][-1]
        if cfg['wiring'] == 'gat':
            ctx_fn = lambda: gen.gatencoder(ctx_in, dev_in['seq_start_end'], end_pos, end_grp)
        else:   # the path TrajectoryGenerator.context() takes: cat + Linear + ReLU + Linear + ReLU in one launch
            from group_gan_gcn_gat_b200 import ops as _ops
            ctx_fn = lambda: _ops.mlp2(gen.mlp_decoder_context, h_enc.view(-1, 32), pool_h)
        ctx_ms = time_call(ctx_fn)
        enc_ms = time_call(lambda: gen.encoder(dev_in['obs_traj_rel']))
        ctx24 = ctx_fn()
        z0 = torch.randn(n_scenes, 8, device=dev)
        dec_ms = time_call(lambda: gen.decode(ctx24, dev_in['obs_traj'], dev_in['obs_traj_rel'], dev_in['seq_start_end'],
                                              user_noise=z0))
        other_modes = {}
        for alt in ('bf16', 'fp32-simt'):                    # the same pooling call in the other modes, for the record
            if alt == pool_mode:
                continue
            gen.pool_net.precision = alt
            other_modes[alt] = time_call(lambda: gen.pool_net(h_enc, dev_in['seq_start_end'], end_pos), reps=5)
        gen.pool_net.precision = precision

    hot_ops = None
    if rank == 0:
        try:
            hot_ops = hot_path_op_numbers(dev, args.scenes, json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json'))).get(
                'hbm_gbs', 6650.0) if os.path.exists(os.path.join(ROOT, 'MEASURED_PEAKS.json')) else 6650.0, flush)
        except Exception as e:                               # the headline must not depend on the sub-measurement
            hot_ops = [{'error': repr(e)[:200]}]

    small = None
    try:
        small = small_batch_numbers(gen, args.config, dev, rank, world)
    except Exception as e:
        small = {'error': repr(e)[:200]}

    train = None
    if not args.no_train:
        try:
            train = train_step_numbers(dev, rank, world)
        except Exception as e:                               # the headline must not depend on the sub-measurement
            train = {'error': repr(e)[:200]}

    t_total = torch.tensor([total_ms, e2e_s * 1e3, pipe_s * 1e3], dtype=torch.float64, device=dev)
    work = torch.tensor([float(peds * K_SAMPLES)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_total, op=dist.ReduceOp.MAX)
        dist.all_reduce(work, op=dist.ReduceOp.SUM)
    total_ms_max, e2e_ms_max, pipe_ms_max = t_total.tolist()
    traj_per_step = work.item()

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
        except OSError:
            pass
        cuda_core_peak = 148 * 128 * 2 * 1.965e-3                      # fp32 FMA lanes x 2 FLOP x GHz -> TFLOP/s
        if pool_mode == 'fp32-simt':
            peak, peak_src = cuda_core_peak, 'fp32 CUDA-core peak (148 SMs x 128 FMA lanes x 1.965 GHz); no fp32 tensor-core MMA exists'
        else:   # fp16 and bf16 operands run at the same tcgen05 rate
            peak = peaks.get('bf16_tflops_sustained', 1400.0)
            peak_src = 'MEASURED_PEAKS.json bf16 sustained (kernel timed inside a long step)' if peaks else 'fallback 1400 (B200_PROFILING.md)'
        k_ms = statistics.mean(kernel_ms)
        achieved = POOL_FLOPS_PER_PAIR * n_pairs / (k_ms * 1e-3) / 1e12
        executed = POOL_EXECUTED_FLOPS[pool_mode] * n_pairs / (k_ms * 1e-3) / 1e12
        traffic, traffic_src = measured_traffic(args.config, kernel, args.scenes)
        h2d = int(sum(host[k].numel() * host[k].element_size() for k in ('obs_traj', 'obs_traj_rel', 'obs_traj_g', 'pred_traj_gt')) +
                  (4 + 4 + 8) * peds + 4 * (n_scenes + 1) + 4 * ((n_pairs + 127) // 128))
        hbm = peaks.get('hbm_gbs', 6650.0)
        if cfg['wiring'] == 'gat':
            ctx_bytes = 260 * peds + 16 * n_scenes + 29920
            ctx_entry = {'op': 'GATEncoder fwd (gat_fused_tc_kernel: tcgen05 linear maps, group structure in-kernel)', 'bound': 'hbm', 'ms': ctx_ms,
                         'algorithmic_bytes': ctx_bytes, 'achieved': ctx_bytes / (ctx_ms * 1e-3) / 1e9, 'peak': hbm,
                         'unit': 'GB/s', 'frac': ctx_bytes / (ctx_ms * 1e-3) / 1e9 / hbm,
                         'algorithmic_flops': 190 * n_pairs + 8400 * peds,
                         'achieved_tflops': (190 * n_pairs + 8400 * peds) / (ctx_ms * 1e-3) / 1e12,
                         'fp32_cuda_core_peak_tflops': cuda_core_peak,
                         'note': 'HBM-bound by decree (SURVEY 8d: 260 B/ped); in practice bound by per-warp instruction latency, DESIGN.md 4.3'}
        else:
            ctx_bytes = (40 + 24) * 4 * peds
            ctx_entry = {'op': 'cat + mlp_decoder_context 40->64->24 (mlp2_tc_kernel: both linear maps on tcgen05)', 'bound': 'hbm', 'ms': ctx_ms,
                         'algorithmic_bytes': ctx_bytes, 'achieved': ctx_bytes / (ctx_ms * 1e-3) / 1e9, 'peak': hbm,
                         'unit': 'GB/s', 'frac': ctx_bytes / (ctx_ms * 1e-3) / 1e9 / hbm}
        line = {
            'metric': 'predicted_trajectories_per_sec', 'value': traj_per_step * args.steps / (total_ms_max * 1e-3),
            'unit': 'traj/s', 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
            'ms_per_step': total_ms_max / args.steps, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': 'bf16' if pool_mode == 'bf16' else 'fp32', 'data': 'synthetic',
            'config': dict(workload_config(args.config, args.scenes), pool_precision=precision),
            'run': {'peds_per_gpu': peds, 'pairs_per_gpu': n_pairs, 'scenes_this_rank': n_scenes, 'pool_kernel': kernel,
                    'pool_mode': pool_mode, 'numa_node_rank0': numa,
                    'pool_mode_note': 'fp32 contract mode (ADE/FDE <= 1e-4, pooled features 1e-5): tcgen05 kind::f16 on '
                                      'fp16 hi+lo operand splits, fp32 accumulate' if pool_mode == 'tc32' else
                                      ('bf16 operands: pooled features 2e-2, ADE/FDE ~1e-3 -- outside the ADE contract'
                                       if pool_mode == 'bf16' else 'fp32 CUDA cores')},
            'clocks': clocks,
            'e2e': {'value': traj_per_step * args.steps / (e2e_ms_max * 1e-3), 'unit': 'traj/s',
                    'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': int(out_host.numel() * 4),
                    'ms_per_step': e2e_ms_max / args.steps, 'ms_per_step_median_rank0': statistics.median(e2e_steps),
                    'ms_per_step_max_rank0': max(e2e_steps),
                    'pipelined': {'value': traj_per_step * args.steps / (pipe_ms_max * 1e-3), 'unit': 'traj/s',
                                  'ms_per_step': pipe_ms_max / args.steps,
                                  'device_ms_per_step_rank0': [round(v, 3) for v in pipe_dev_ms],
                                  'host_issue_ms_per_step_rank0': [round(v, 3) for v in pipe_host_ms],
                                  'what': 'same steps, next batch copied on a side stream during compute, async D2H per '
                                          'step, one synchronisation at the end (evaluate() over a prefetching loader)'},
                    'what': 'evaluate_batch(): H2D batch + schedule + K forwards + best-of-K ADE/FDE on device + D2H of the sums'},
            'gpu_launches': int(launches),
            'roofline': {'kernel': kernel, 'bound': 'tensor' if pool_mode != 'fp32-simt' else 'fp32 cuda cores',
                         'achieved': achieved, 'peak': peak, 'unit': 'TFLOP/s', 'frac': achieved / peak,
                         'executed_tflops': executed, 'frac_executed': executed / peak,
                         'traffic': traffic, 'traffic_source': traffic_src,
                         'peak_source': peak_src, 'kernel_ms': k_ms,
                         'share_of_step': k_ms * K_SAMPLES / (total_ms_max / args.steps),
                         'algorithmic_flops_per_launch': POOL_FLOPS_PER_PAIR * n_pairs,
                         'executed_flops_per_launch': POOL_EXECUTED_FLOPS[pool_mode] * n_pairs,
                         'note': 'achieved/frac: as-written FLOPs (57408 per ordered pair, SURVEY 8d); executed: the MMA FLOPs '
                                 'the kernel issues (tc32: K = 16 + 3 x 32 for GEMM1, hidden hi+lo against [W2_hi; W2_lo] '
                                 'for GEMM2; the N = 16 GEMM2 MMAs are A-operand-fetch bound at 32 cycles each, DESIGN.md 4.1b)',
                         'same_call_other_modes_ms': other_modes, 'op_ms_with_prep_and_unpack': pool_ms},
            'wall_s_timed_region': wall,
            'other_kernels': [
                ctx_entry,
                {'op': 'Encoder + decoder LSTM', 'bound': 'xu (MUFU)', 'ms': enc_ms + dec_ms,
                 'mufu_per_ped_step': 7 * 32,
                 'achieved_gmufu_s': 224 * (OBS_LEN + pred_len) * peds / ((enc_ms + dec_ms) * 1e-3) / 1e9,
                 'peak_gmufu_s': 148 * 16 * 1.965,
                 'frac': 224 * (OBS_LEN + pred_len) * peds / ((enc_ms + dec_ms) * 1e-3) / 1e9 / (148 * 16 * 1.965),
                 'note': '7 transcendental operations per hidden unit and step on 16 MUFU lanes per SM at 1.965 GHz (DESIGN.md 4.6)'},
                {'op': 'Encoder LSTM 8 steps (lstm_tc_kernel)', 'ms': enc_ms, 'ped_steps_per_s': OBS_LEN * peds / (enc_ms * 1e-3)},
                {'op': 'Decoder LSTM %d steps + hidden2pos + noise fold-in (lstm_tc_kernel)' % pred_len, 'ms': dec_ms,
                 'ped_steps_per_s': pred_len * peds / (dec_ms * 1e-3)},
            ],
        }
        if hot_ops is not None:
            line['hot_path_ops'] = hot_ops
        if world > 1:
            line['nccl'] = nccl_evidence(nccl_log)
        if small is not None:
            line['small_batch'] = small
        if train is not None:
            line['train_step'] = train
        if not args.no_cpu_baseline and world == 1:          # the CPU port is timed next to the 1-GPU number only
            v, dt, p = cpu_port_traj_per_sec(8192, K_SAMPLES, 1234 + 2, config=args.config)   # ~10-20 s of CPU work
            line['cpu_baseline'] = {'value': v, 'unit': 'traj/s', 'cores': os.cpu_count(), 'kind': 'port',
                                    'sample': '8192 scenes of the config\'s histogram (%d peds) x K=20 forwards, %.1f s' % (p, dt)}
    if world > 1:
        dist.destroy_process_group()
    if rank == 0:
        # after the process group is gone and the other ranks had time to finish writing: a line > 4 KB is not written
        # atomically to a pipe the ranks share
        if world > 1:
            time.sleep(0.5)
        sys.stdout.flush()
        print(json.dumps(line), flush=True)


def hot_path_op_numbers(dev, n_scenes, hbm, flush):
    """The other hot-path operators of north_star next to the pooling kernel, op-level (module call, CUDA events, L2
    flushed), on zara1-shaped scenes of the bench size: GATEncoder / GCNModule forward (tcgen05 kernels) and
    forward+backward (single-launch backward kernels), PoolHiddenNet backward (argmax-sparse).  Algorithmic bytes per
    pedestrian: SURVEY 8d (260 B forward, 420 B backward for the graph operators)."""
    from group_gan_gcn_gat_b200 import modules as M
    data = synth_batch(n_scenes, 1237, 'sgan_gat')
    sse = data['seq_start_end'].to(dev)
    n = int(sse[-1, 1])
    lab, pos = data['obs_traj_g'][-1].to(dev), data['obs_traj'][-1].to(dev)

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
    rows = []
    torch.manual_seed(0)
    for name, mod, kern in (('GATEncoder', M.GATEncoder(None, 1, 0, 0.2), 'gat_fused_tc_kernel'),
                            ('GCNModule', M.GCNModule(), 'gcn_fused_tc_kernel')):
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
        # the backward operator alone (no autograd engine around it: the module-level number above carries ~0.3 ms of
        # host time -- engine thread hand-over, ten gradient allocations -- that the GPU waits for at this size)
        from group_gan_gcn_gat_b200 import ops
        from group_gan_gcn_gat_b200.modules import _groups_for
        from group_gan_gcn_gat_b200.schedule import get_schedule
        sched = get_schedule(sse, dev)
        leader, gsize, _gid, ngrp = _groups_for(sched, lab)
        chunk_scene, n_chunks = sched.chunks(32)
        with torch.no_grad():
            if name == 'GATEncoder':
                ws = (*mod.gat_intra.stacked(), *mod.gat_inter.stacked(), mod.out_embedding.weight, mod.out_embedding.bias)
                b_only = timed(lambda: ops._DIRECT[ops.gat_encoder_bwd](
                    x, up, leader, gsize, sched.ped_start, sched.ped_end, sched.n_scenes, *ws, 0.2, sched.scene_start,
                    chunk_scene, n_chunks, 32, int(sched.max_n)))
            else:
                ws = (mod.gcn_intra.W[0], mod.gcn_intra.W[1], mod.gcn_inter.W[0], mod.gcn_inter.W[1],
                      mod.out_embedding.weight, mod.out_embedding.bias)
                b_only = timed(lambda: ops._DIRECT[ops.gcn_module_bwd](
                    x, up, leader, gsize, sched.ped_start, sched.ped_end, sched.scene_start, ngrp, *ws, chunk_scene, n_chunks))
        fwd_b, bwd_b = 260 * n + 16 * n_scenes, 420 * n
        rows.append({'op': name + ' fwd (%s, group structure in-kernel)' % kern, 'bound': 'hbm', 'peds': n, 'ms': f_ms,
                     'algorithmic_bytes': fwd_b, 'achieved': fwd_b / f_ms / 1e6, 'peak': hbm, 'unit': 'GB/s',
                     'frac': fwd_b / f_ms / 1e6 / hbm})
        rows.append({'op': name + ' bwd (operator call: single-launch backward kernel + reduction)', 'bound': 'hbm', 'peds': n,
                     'ms': b_only, 'algorithmic_bytes': bwd_b, 'achieved': bwd_b / b_only / 1e6, 'peak': hbm, 'unit': 'GB/s',
                     'frac': bwd_b / b_only / 1e6 / hbm})
        rows.append({'op': name + ' fwd + bwd (autograd: forward, group_ids, single-launch backward, reduction)',
                     'bound': 'hbm', 'peds': n, 'ms': fb_ms, 'algorithmic_bytes': fwd_b + bwd_b,
                     'achieved': (fwd_b + bwd_b) / fb_ms / 1e6, 'peak': hbm, 'unit': 'GB/s',
                     'frac': (fwd_b + bwd_b) / fb_ms / 1e6 / hbm})
    # PoolHiddenNet backward: gradients flow through the argmax pair of every (pedestrian, channel) only
    pool = M.PoolHiddenNet(embedding_dim=16, h_dim=32, mlp_dim=64, bottleneck_dim=8, batch_norm=False).to(dev)
    h = torch.randn(n, 32, device=dev, requires_grad=True)
    out = pool(h, sse, pos)
    go = torch.randn_like(out)
    params = [h] + list(pool.parameters())
    b_ms = timed(lambda: torch.autograd.grad(out, params, go, retain_graph=True))
    pool_bwd_b = n * (8 * 4 + 8 * 4 + 2 * 4 + 2 * 32 * 4)      # grad_out, argmax, position, h read + grad_h written
    rows.append({'op': 'PoolHiddenNet bwd (sgx_pool_bwd_scenes: scene-owned event kernel + 3 tensor-core GEMMs over the [batch,512] hidden gradient; argmax-sparse, 8 pairs per pedestrian)', 'bound': 'hbm', 'peds': n,
                 'ms': b_ms, 'algorithmic_bytes': pool_bwd_b, 'achieved': pool_bwd_b / b_ms / 1e6, 'peak': hbm,
                 'unit': 'GB/s', 'frac': pool_bwd_b / b_ms / 1e6 / hbm})
    return rows


def small_batch_numbers(gen, config, dev, rank, world, scenes=64, reps=20):
    """ONE minibatch at the reference's own batch size (64 scenes, scripts/evaluate_model.py:72-99), best-of-20: the
    launch-latency regime where scene sharding runs dry.  N = 1: the K samples folded into one forward; N > 1: the
    (sample, scene) pairs sharded over the ranks by LPT + one all-reduce of the [K, S] error sums."""
    import torch.distributed as dist
    from group_gan_gcn_gat_b200 import parallel
    from group_gan_gcn_gat_b200.evaluate import evaluate_batch
    data = synth_batch(scenes, 4242, config)
    d = {k: data[k].to(dev) for k in ('obs_traj', 'obs_traj_rel', 'obs_traj_g', 'pred_traj_gt')}
    sse = data['seq_start_end']
    noise = torch.randn(K_SAMPLES, scenes, 8, generator=torch.Generator().manual_seed(11)).to(dev)
    peds = int(sse[-1, 1])

    def run():
        if world == 1:
            return evaluate_batch(gen, d['obs_traj'], d['obs_traj_rel'], sse, d['obs_traj_g'], d['pred_traj_gt'], K_SAMPLES,
                                  noise=noise, fold_samples=True)
        return parallel.evaluate_batch_sample_sharded(gen, d['obs_traj'], d['obs_traj_rel'], sse, d['obs_traj_g'],
                                                      d['pred_traj_gt'], K_SAMPLES, noise, world, rank)
    ts = []
    with torch.no_grad():
        for i in range(reps + 5):
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            a, f = run()
            torch.cuda.synchronize()
            ts.append((time.perf_counter() - t0) * 1e3)
    t = torch.tensor([statistics.median(ts[5:])], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    return {'workload': 'one %d-scene minibatch (%d peds), best-of-%d, wall clock incl. host issue, median of %d, max over ranks'
                        % (scenes, peds, K_SAMPLES, reps),
            'mode': 'K samples folded into one forward' if world == 1 else
                    '(sample, scene) pairs LPT-sharded over %d ranks + all-reduce of the [K,S] sums' % world,
            'ms': ms, 'traj_per_s': peds * K_SAMPLES / (ms * 1e-3), 'ade_sum': float(a), 'fde_sum': float(f)}


def run_train_mode(args, dev, rank, world, local_rank, nccl_log, L):
    """bench.py --mode train: SURVEY cfg 5 as the line's own metric.  One step = one adversarial iteration (discriminator
    step + generator step with best_k = 20, Adam, gradient all-reduce over NCCL) on 64 zara1-train-shaped scenes per rank
    (weak scaling).  value: batch resident in HBM; e2e: the batch copied from pinned host memory every step and the
    losses read back."""
    import torch.distributed as dist
    from tools import train_step_dp as T
    from group_gan_gcn_gat_b200 import parallel
    targs, gen, disc, opt_g, opt_d, batch = T.setup(dev, rank, world)
    host = tuple(t.cpu().pin_memory() for t in batch)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step(it, b):
        rng = parallel.make_label_rng(0, it)
        torch.manual_seed(100 + it * world + rank)
        parallel.discriminator_step(targs, b, gen, disc, opt_d, label_rng=rng, n_global=targs.n_global)
        return parallel.generator_step(targs, b, gen, disc, opt_g, label_rng=rng, n_global=targs.n_global)

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    for it in range(args.warmup):
        step(it, batch)
    barrier()
    launches0 = L.sgx_launch_count()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    wall0 = time.perf_counter()
    for it, (a, b) in enumerate(ev):
        a.record()
        step(args.warmup + it, batch)
        b.record()
    barrier()
    wall = time.perf_counter() - wall0
    launches = L.sgx_launch_count() - launches0
    if rank == 0:
        sampler.window(wall0, wall0 + wall)
    clocks = sampler.stop() if rank == 0 else None
    total_ms = sum(a.elapsed_time(b) for a, b in ev)
    out_host = torch.empty(1).pin_memory()
    barrier()
    t0 = time.perf_counter()
    for it in range(args.steps):
        b = tuple(h.to(dev, non_blocking=True) for h in host)
        losses = step(args.warmup + args.steps + it, b)
        out_host.copy_(torch.as_tensor(losses['G_total_loss'], device=dev).reshape(1), non_blocking=True)
        torch.cuda.synchronize()
    barrier()
    e2e_ms = (time.perf_counter() - t0) * 1e3
    t = torch.tensor([total_ms, e2e_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms, e2e_ms = t.tolist()
    if rank == 0:
        scenes = 64 * world
        line = {'metric': 'training_scenes_per_sec', 'value': scenes * args.steps / (total_ms * 1e-3), 'unit': 'scenes/s',
                'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': total_ms / args.steps,
                'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'fp32', 'data': 'synthetic',
                'config': {'workload': 'SURVEY cfg 5: adversarial step (D step + G step, best_k = 20, Adam) of SGAN-GAT G + '
                                       'global pooled D, 64 zara1-train-shaped scenes per rank, gradient all-reduce',
                           'scenes_per_gpu': 64, 'peds_global': int(targs.n_global), 'best_k': 20,
                           'l2': 'working set (< 10 MB) is L2 resident by nature of the workload: one minibatch',
                           'parallelism': 'scenes sharded by LPT on N^2, NCCL all-reduce of G and D gradients'},
                'clocks': clocks,
                'e2e': {'value': scenes * args.steps / (e2e_ms * 1e-3), 'unit': 'scenes/s',
                        'h2d_bytes_per_step': int(sum(h.numel() * h.element_size() for h in host)), 'd2h_bytes_per_step': 4,
                        'ms_per_step': e2e_ms / args.steps},
                'gpu_launches': int(launches),
                'roofline': {'kernel': 'launch-latency bound step (hundreds of launches of < 10 us on 800 pedestrians)',
                             'bound': 'hbm', 'achieved': None, 'peak': None, 'unit': 'GB/s', 'frac': None, 'traffic': None},
                'allreduce_bytes': {'G': sum(p.numel() for p in gen.parameters()) * 4,
                                    'D': sum(p.numel() for p in disc.parameters()) * 4}}
        if world > 1:
            line['nccl'] = nccl_evidence(nccl_log)
        print(json.dumps(line), flush=True)


def train_step_numbers(dev, rank, world):
    """SURVEY cfg 5 next to the headline: one adversarial iteration (D step + G step, best_k = 20) on 64 zara1-train-shaped
    scenes per rank, data-parallel with the gradient all-reduce of group_gan_gcn_gat_b200.parallel; device time per
    step, max over ranks.  The driver's scaling run therefore also records the training step at N = 1, 2, 4, 8."""
    import torch.distributed as dist
    from tools import train_step_dp as T
    from group_gan_gcn_gat_b200 import parallel
    args, gen, disc, opt_g, opt_d, batch = T.setup(dev, rank, world)
    times = {'d': [], 'g': []}
    for it in range(6):
        rng = parallel.make_label_rng(0, it)
        torch.manual_seed(100 + it * world + rank)
        for name, fn, opt in (('d', parallel.discriminator_step, opt_d), ('g', parallel.generator_step, opt_g)):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            e0.record()
            fn(args, batch, gen, disc, opt, label_rng=rng, n_global=args.n_global)
            e1.record()
            torch.cuda.synchronize()
            times[name].append(e0.elapsed_time(e1))
    t = torch.tensor([statistics.median(times['d'][2:]), statistics.median(times['g'][2:])], device=dev, dtype=torch.float64)
    n = torch.tensor([float(batch[0].shape[1])], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(n, op=dist.ReduceOp.SUM)
    d_ms, g_ms = t.tolist()
    return {'workload': 'cfg5 adversarial step: 64 zara1-train-shaped scenes per rank, best_k=20, SGAN-GAT G + global D, Adam',
            'd_step_ms': d_ms, 'g_step_ms': g_ms, 'peds_global': int(n.item()), 'scenes_global': 64 * world,
            'scenes_per_s': 64 * world / ((d_ms + g_ms) * 1e-3),
            'allreduce_bytes': {'G': sum(p.numel() for p in gen.parameters()) * 4,
                                'D': sum(p.numel() for p in disc.parameters()) * 4}}


if __name__ == '__main__':
    main()
