"""Best-of-K evaluation of one minibatch -- the loop body of scripts/evaluate_model.py:72-99 without per-scene
python loops (SURVEY 8f rows f2/f3 stay behind the same API: every sample is a complete generator forward).

    ade_sum, fde_sum = evaluate_batch(generator, obs_traj, obs_traj_rel, seq_start_end, obs_traj_g, pred_traj_gt, K)

ade_sum / (total_traj * pred_len) and fde_sum / total_traj are the numbers the reference prints.
"""
import torch

from .losses import displacement_error, final_displacement_error
from .models import ped_scene_index
from .schedule import get_schedule, tiled_schedule
from .utils import ready, ready_last, relative_to_abs, stage_host_batch


def best_of_k_sum(per_sample, sched):
    """evaluate_helper (scripts/evaluate_model.py:58-69): per scene sum over peds, min over the K samples, summed."""
    seg = ped_scene_index(sched)
    per_scene = per_sample.new_zeros(sched.n_scenes, per_sample.shape[1]).index_add_(0, seg, per_sample)
    return per_scene.min(dim=1).values.sum()


@torch.no_grad()
def evaluate_batch(generator, obs_traj, obs_traj_rel, seq_start_end, obs_traj_g, pred_traj_gt, num_samples=20,
                   noise=None, hoist_context=False, fold_samples='auto'):
    """fold_samples: run the num_samples forwards as ONE forward over num_samples copies of the batch (they share
    weights and inputs and differ only in the noise).  At the reference's own batch size (64 scenes, ~250 peds) the
    sample loop is launch-latency bound: 8.0 ms for 20 forwards against 1.2 ms folded.  'auto' folds when the folded
    batch stays below 2^20 pedestrians (above that a single forward already fills the GPU) and nothing in the
    generator couples batch rows (BatchNorm, active dropout).
    hoist_context=True computes the noise-independent part of the forward (encoder, pooling, graph context:
    everything before sgan/models.py:909) once instead of num_samples times -- bit-identical results when the
    decoder does not pool per step.  Off by default: the reference recomputes it per sample.
    HOST tensors (scripts/evaluate_model.py:75 receives the batch from a CPU DataLoader) are staged here: copied on a copy
    stream in the order the forward reads them, the compute stream waiting for each where it is first needed
    (utils.stage_host_batch), so the first sample's encoder overlaps most of the transfer; pin them for that."""
    if not obs_traj.is_cuda:
        gdev = next(generator.parameters()).device
        if gdev.type == 'cuda':
            obs_traj, obs_traj_rel, obs_traj_g, pred_traj_gt = stage_host_batch(gdev, obs_traj, obs_traj_rel, obs_traj_g,
                                                                                pred_traj_gt)
            if noise is not None:
                noise = noise.to(gdev, non_blocking=True)
    # The schedule is built where it is first needed (the pooling call of the first forward, the folded copies, the
    # best-of-K reduction) and cached on seq_start_end: the first sample's encoder is queued before the host touches it.
    n_scenes = seq_start_end.n_scenes if hasattr(seq_start_end, 'n_scenes') else int(seq_start_end.shape[0])
    if noise is None and generator.noise_dim and generator.noise_mix_type == 'global':
        # one draw for all K samples on the device generator.  The reference draws each sample on the CPU generator and
        # copies it (sgan/models.py:23-29), which costs more host time than the whole forward here; pass `noise=` to
        # reproduce a CPU-seeded stream.
        fn = torch.randn if generator.noise_type == 'gaussian' else (lambda *a, **k: torch.rand(*a, **k) * 2 - 1)
        noise = fn(num_samples, n_scenes, *generator.noise_dim, device=obs_traj.device)
    dev = obs_traj.device
    if fold_samples == 'auto':
        fold_samples = num_samples > 1 and obs_traj.shape[1] * num_samples <= (1 << 20) and not hoist_context
    if fold_samples and obs_traj.is_cuda and num_samples <= 32 and generator.noise_mix_type == 'global' and \
            noise is not None and _batch_independent(generator):
        return _evaluate_batch_folded(generator, obs_traj, obs_traj_rel, seq_start_end, obs_traj_g, pred_traj_gt,
                                      num_samples, noise, get_schedule(seq_start_end, obs_traj.device))
    if obs_traj.is_cuda and num_samples <= 32:
        # fused path: one metrics kernel per sample, one best-of-K kernel (sgx_displacement_errors / sgx_best_of_k)
        from . import _lib
        from .ops import _f32, _ptr, _stream
        L = _lib.lib()
        batch, T = obs_traj.shape[1], pred_traj_gt.shape[0]
        ade = torch.empty(batch, num_samples, dtype=torch.float32, device=dev)
        fde = torch.empty_like(ade)
        out2 = torch.empty(2, dtype=torch.float32, device=dev)
        ctx = generator.context(obs_traj, obs_traj_rel, seq_start_end, obs_traj_g) if hoist_context else None
        gt = start = None
        with torch.cuda.device(dev):
            for k in range(num_samples):
                nz = None if noise is None else noise[k]
                if ctx is not None:
                    rel = generator.decode(ctx, obs_traj, obs_traj_rel, seq_start_end, user_noise=nz).contiguous()
                else:
                    rel = generator(obs_traj, obs_traj_rel, seq_start_end, obs_traj_g, user_noise=nz).contiguous()
                if gt is None:       # after the first forward is queued: a staged pred_traj_gt is the last tensor to arrive
                    gt, start = _f32(ready(pred_traj_gt), 'pred_traj_gt'), _f32(ready_last(obs_traj)[-1], 'obs_traj')
                _lib.check(L.sgx_displacement_errors(_ptr(rel), _ptr(start), _ptr(gt), T, batch, _ptr(ade), _ptr(fde),
                                                     num_samples, k, _stream(rel)), 'sgx_displacement_errors')
            sched = get_schedule(seq_start_end, dev)                       # (cached by the forwards above)
            _lib.check(L.sgx_best_of_k(_ptr(ade), _ptr(fde), _ptr(sched.scene_start), sched.n_scenes, num_samples,
                                       _ptr(out2), _stream(ade)), 'sgx_best_of_k')
        return out2[0], out2[1]
    ade, fde = [], []
    ready_last(obs_traj), ready(pred_traj_gt)
    for k in range(num_samples):
        rel = generator(obs_traj, obs_traj_rel, seq_start_end, obs_traj_g,
                        user_noise=None if noise is None else noise[k])
        pred = relative_to_abs(rel, obs_traj[-1])
        ade.append(displacement_error(pred, pred_traj_gt, mode='raw'))
        fde.append(final_displacement_error(pred[-1], pred_traj_gt[-1], mode='raw'))
    sched = get_schedule(seq_start_end, obs_traj.device)
    return best_of_k_sum(torch.stack(ade, dim=1), sched), best_of_k_sum(torch.stack(fde, dim=1), sched)


def _batch_independent(module):
    from .parallel import _batch_independent as f
    return f(module)


def _evaluate_batch_folded(generator, obs_traj, obs_traj_rel, seq_start_end, obs_traj_g, pred_traj_gt, k, noise, sched):
    """k samples as one forward over k copies of the batch (sample-major), one metrics launch, one best-of-k launch."""
    from . import _lib
    from .ops import _f32, _ptr, _stream
    L = _lib.lib()
    dev = obs_traj.device
    for t in (obs_traj, obs_traj_rel, obs_traj_g, pred_traj_gt):
        ready(t)                                           # (staged host batch: the folded copies read all of them at once)
    n, s, T = obs_traj.shape[1], sched.n_scenes, pred_traj_gt.shape[0]
    sse_k = tiled_schedule(sched, k, dev)                # built on the host from the base schedule: no device read-back
    rel = generator(obs_traj.repeat(1, k, 1), obs_traj_rel.repeat(1, k, 1), sse_k, obs_traj_g.repeat(1, k, 1),
                    user_noise=noise.reshape(k * s, -1)).contiguous()
    gt = _f32(pred_traj_gt.repeat(1, k, 1), 'pred_traj_gt')
    start = _f32(obs_traj[-1].repeat(k, 1), 'obs_traj')
    ade = torch.empty(k * n, 1, dtype=torch.float32, device=dev)
    fde = torch.empty_like(ade)
    out2 = torch.empty(2, dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        _lib.check(L.sgx_displacement_errors(_ptr(rel), _ptr(start), _ptr(gt), T, k * n, _ptr(ade), _ptr(fde), 1, 0,
                                             _stream(rel)), 'sgx_displacement_errors')
        ade_nk = ade.view(k, n).t().contiguous()           # [n, k]: column = sample, as the per-sample path writes it
        fde_nk = fde.view(k, n).t().contiguous()
        _lib.check(L.sgx_best_of_k(_ptr(ade_nk), _ptr(fde_nk), _ptr(sched.scene_start), s, k, _ptr(out2), _stream(ade)),
                   'sgx_best_of_k')
    return out2[0], out2[1]


def get_generator(checkpoint, device='cuda', context_type='gat', pool_precision=None):
    """scripts/evaluate_model.py:26-55: a TrajectoryGenerator built from checkpoint['args'] with checkpoint['g_state']
    loaded, on `device`, in train mode (:54).  n_units = [40] + hidden_units + [40] as at :23-27 (GATEncoder ignores
    it); checkpoints written before the GAT arguments existed (sgan-p / sgan-g models) get hidden_units '16',
    n_heads 1, dropout1 0, alpha 0.2 -- pass context_type='mlp' / 'gcn' for those wirings."""
    from .models import TrajectoryGenerator
    args = checkpoint['args']
    get = (lambda k, d=None: args.get(k, d)) if isinstance(args, dict) else (lambda k, d=None: getattr(args, k, d))
    hidden = [int(x) for x in str(get('hidden_units', '16')).strip().split(',')]
    n_heads = get('n_heads', 1)
    n_heads = n_heads if isinstance(n_heads, int) else int(str(n_heads).strip().split(',')[0])
    n_units = [40] + hidden + [40]
    generator = TrajectoryGenerator(
        obs_len=get('obs_len'), pred_len=get('pred_len'), embedding_dim=get('embedding_dim'),
        encoder_h_dim=get('encoder_h_dim_g'), decoder_h_dim=get('decoder_h_dim_g'), mlp_dim=get('mlp_dim'),
        num_layers=get('num_layers'), noise_dim=tuple(get('noise_dim')), noise_type=get('noise_type'),
        noise_mix_type=get('noise_mix_type'), pooling_type=get('pooling_type'),
        pool_every_timestep=get('pool_every_timestep'), dropout=get('dropout'), bottleneck_dim=get('bottleneck_dim'),
        neighborhood_size=get('neighborhood_size'), grid_size=get('grid_size'), batch_norm=get('batch_norm'),
        n_units=n_units, n_heads=n_heads, dropout1=get('dropout1', 0), alpha=get('alpha', 0.2),
        context_type=context_type)
    generator.load_state_dict(checkpoint['g_state'])
    generator = generator.to(device).train()
    if pool_precision is not None and getattr(generator, 'pool_net', None) is not None:
        generator.pool_net.precision = pool_precision
    return generator


@torch.no_grad()
def evaluate(args, loader, generator, num_samples, noise_for_batch=None, hoist_context=False, fold_samples='auto'):
    """scripts/evaluate_model.py:72-99: best-of-`num_samples` ADE / FDE over a loader of 11-tuples (data.seq_collate /
    data.DeviceLoader).  Batches already on the generator's device are used as they are; host batches are staged by
    evaluate_batch (copy stream, in the order the forward reads them).
    `noise_for_batch(batch_index, n_scenes) -> [num_samples, n_scenes, *noise_dim]` pins the noise (parity runs)."""
    device = next(generator.parameters()).device
    ade_sum = torch.zeros((), dtype=torch.float64, device=device)
    fde_sum = torch.zeros((), dtype=torch.float64, device=device)
    total_traj = 0
    for b, batch in enumerate(loader):
        # host batches stay on the host here: evaluate_batch stages them in the order the forward reads them
        obs_traj, pred_traj_gt, obs_traj_rel = batch[0], batch[1], batch[2]
        obs_traj_g, seq_start_end = batch[6], batch[10]
        total_traj += pred_traj_gt.size(1)
        noise = None if noise_for_batch is None else noise_for_batch(b, seq_start_end.size(0)).to(device)
        a, f = evaluate_batch(generator, obs_traj.contiguous(), obs_traj_rel.contiguous(), seq_start_end,
                              obs_traj_g.contiguous(), pred_traj_gt.contiguous(), num_samples, noise=noise,
                              hoist_context=hoist_context, fold_samples=fold_samples)
        ade_sum += a.double()
        fde_sum += f.double()
    pred_len = args['pred_len'] if isinstance(args, dict) else args.pred_len
    ade = ade_sum / (total_traj * pred_len)
    fde = fde_sum / total_traj
    return ade.float(), fde.float()


class GraphedGenerator:
    """Replays generator(obs_traj, obs_traj_rel, seq_start_end, obs_traj_g, user_noise) from a CUDA graph.

    The launch-latency regime (a 64-scene minibatch is ~250 pedestrians: 13 kernels of a few microseconds each) is bound
    by host issue time; every sgx entry point is capture-safe (no allocation, no synchronisation, the caller's stream),
    so one forward is captured per scene layout and replayed with new trajectories / noise copied into its static
    buffers: 0.93 -> 0.45 ms for the best-of-20 forward of a 64-scene batch.  A different `seq_start_end` (different
    ragged layout = different grids) triggers a re-capture; at most `max_graphs` layouts are kept.
    Inference only (no autograd through a graph replay).  The captured launches read the prepared weight images the modules
    cache per weight version (pooling, recurrences, graph context), so a graph must be re-captured after the weights change
    (`GraphedGenerator(generator)` again, or `.reset()`).
    """

    def __init__(self, generator, max_graphs=8):
        self.generator, self.max_graphs, self._graphs = generator, max_graphs, {}

    def reset(self):
        """drop the captured graphs (after a weight update)"""
        self._graphs.clear()

    def _capture(self, key, obs_traj, obs_traj_rel, seq_start_end, obs_traj_g, user_noise):
        static = [t.clone() for t in (obs_traj, obs_traj_rel, obs_traj_g, user_noise)]
        sse = seq_start_end.clone()
        side = torch.cuda.Stream(obs_traj.device)
        side.wait_stream(torch.cuda.current_stream(obs_traj.device))
        with torch.cuda.stream(side), torch.no_grad():            # warm-up: workspaces, schedule and lazy inits
            for _ in range(2):
                self.generator(static[0], static[1], sse, static[2], user_noise=static[3])
        torch.cuda.current_stream(obs_traj.device).wait_stream(side)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph), torch.no_grad():
            out = self.generator(static[0], static[1], sse, static[2], user_noise=static[3])
        if len(self._graphs) >= self.max_graphs:
            self._graphs.pop(next(iter(self._graphs)))
        self._graphs[key] = (graph, static, sse, out)
        return self._graphs[key]

    @torch.no_grad()
    def __call__(self, obs_traj, obs_traj_rel, seq_start_end, obs_traj_g, user_noise):
        if user_noise is None:
            raise ValueError('GraphedGenerator needs user_noise: a draw inside the captured forward would be frozen')
        key = (tuple(obs_traj.shape), tuple(user_noise.shape), seq_start_end.cpu().numpy().tobytes())
        hit = self._graphs.get(key) or self._capture(key, obs_traj, obs_traj_rel, seq_start_end, obs_traj_g, user_noise)
        graph, static, _sse, out = hit
        for dst, src in zip(static, (obs_traj, obs_traj_rel, obs_traj_g, user_noise)):
            dst.copy_(src, non_blocking=True)
        graph.replay()
        return out
