"""Scene-sharded data parallelism and the adversarial step of scripts/train.py:395-484.

Inference shards: scenes are independent units (all three operators are block-diagonal over scenes, the K samples
are independent given the noise), so `shard_batch` splits them across ranks by LPT on N^2 -- computed identically
on every rank from the host copy of seq_start_end, no communication -- and nothing is exchanged in the forward.
Training: one flattened all-reduce (sum) per network per optimizer step; BCE terms are means over ALL pedestrians of
the global batch (sgan/losses.py:21) so the local term is weighted by local_peds / global_peds, the best-of-K L2
variety term is a SUM over scenes (scripts/train.py:460-466) so local sums simply add; the label-smoothing draws
come from an identically seeded `random.Random` on every rank; clip_grad_norm_ acts on the reduced gradient.
"""
import random

import numpy as np
import torch
import torch.distributed as dist
import torch.nn as nn

from .losses import gan_d_loss, gan_g_loss, l2_loss
from .models import ped_scene_index
from .schedule import SceneSchedule, get_schedule, tiled_schedule
from .utils import relative_to_abs


def shard_batch(batch_tensors, seq_start_end, world, rank, ped_dim=1):
    """Returns (local tensors, local seq_start_end [int64, host], scene indices) for `rank`.

    batch_tensors: dict name -> tensor whose dimension `ped_dim` indexes pedestrians ([T,batch,C] in the reference),
    or dimension 0 for tensors listed with a leading '0:' in the name (e.g. loss_mask [batch, T]).
    A rank can come out EMPTY (fewer scenes than ranks: the last minibatch of an epoch, the reference loader has no
    drop_last): its tensors have zero pedestrians and its seq_start_end zero rows; the step functions below skip the
    forward on such a rank but still join every collective.
    """
    sched = seq_start_end if isinstance(seq_start_end, SceneSchedule) else SceneSchedule(seq_start_end, 'cpu')
    rank_of, _ = sched.partition(world)
    mine = np.nonzero(rank_of == rank)[0]
    sse = sched.host_sse
    sizes = (sse[mine, 1] - sse[mine, 0]).astype(np.int64)
    idx = np.concatenate([np.arange(sse[s, 0], sse[s, 1]) for s in mine]) if len(mine) else np.zeros(0, np.int64)
    starts = np.concatenate([[0], np.cumsum(sizes)])
    local_sse = torch.from_numpy(np.stack([starts[:-1], starts[1:]], axis=1).astype(np.int64))
    out = {}
    for name, t in batch_tensors.items():
        dim = 0 if name.startswith('0:') else ped_dim
        sel = torch.from_numpy(idx).to(t.device)
        out[name[2:] if name.startswith('0:') else name] = t.index_select(dim, sel)
    return out, local_sse, mine


def global_ped_count(seq_start_end):
    """Pedestrians of the GLOBAL minibatch, known on every rank without communication (each rank holds the host copy of
    the global seq_start_end it sharded from): pass it as `n_global` to the step functions."""
    sse = seq_start_end.host_sse if isinstance(seq_start_end, SceneSchedule) else np.asarray(seq_start_end)
    return int(sse[-1, 1]) if len(sse) else 0


def partition_samples(seq_start_end, num_samples, world):
    """LPT split of the (sample k, scene s) PAIRS of one minibatch over `world` ranks by cost N_s^2 ("by
    scenes ... and by the K = 20 best-of-K samples"): where one 64-scene minibatch cannot feed 8 GPUs by scenes alone,
    its 20 x 64 (sample, scene) pairs can.  -> rank_of_pair int32 [num_samples, S]; identical on every rank (computed from
    the host copy of seq_start_end, no communication)."""
    sched = seq_start_end if isinstance(seq_start_end, SceneSchedule) else SceneSchedule(seq_start_end, 'cpu')
    sse = sched.host_sse
    sizes = (sse[:, 1] - sse[:, 0]).astype(np.int64)
    tiled = np.tile(sizes, num_samples)                      # virtual scene list: pair (k, s) = scene k * S + s
    ends = np.cumsum(tiled)
    virt = np.stack([ends - tiled, ends], axis=1).astype(np.int64)
    rank_of, _ = SceneSchedule(virt, 'cpu').partition(world)
    return rank_of.reshape(num_samples, len(sizes))


def _sample_shard_plan(sched, K, world, rank, dev):
    """This rank's (sample, scene) pairs of a minibatch as gather indices, cached on the schedule (one minibatch layout is
    evaluated many times: the plan costs more host time than the forward it feeds).  None for an empty shard."""
    cache = sched.__dict__.setdefault('_sample_shards', {})
    key = (K, world, rank, str(dev))
    if key not in cache:
        sse = sched.host_sse
        S = sched.n_scenes
        mine_k, mine_s = np.nonzero(partition_samples(sched, K, world) == rank)      # sample-major order
        if not len(mine_k):
            cache[key] = None
        else:
            sizes = (sse[mine_s, 1] - sse[mine_s, 0]).astype(np.int64)
            idx = torch.from_numpy(np.concatenate([np.arange(sse[s, 0], sse[s, 1]) for s in mine_s])).to(dev)
            ends = np.cumsum(sizes)
            sse_l = torch.from_numpy(np.stack([ends - sizes, ends], axis=1).astype(np.int64))
            pair_of_ped = torch.from_numpy(np.repeat(mine_k * S + mine_s, sizes)).to(dev)
            cache[key] = (idx, sse_l, torch.from_numpy(mine_k).to(dev), torch.from_numpy(mine_s).to(dev), pair_of_ped)
    return cache[key]


@torch.no_grad()
def evaluate_batch_sample_sharded(generator, obs_traj, obs_traj_rel, seq_start_end, obs_traj_g, pred_traj_gt, num_samples,
                                  noise, world=None, rank=None, group=None):
    """Best-of-K ADE / FDE sums of ONE minibatch (scripts/evaluate_model.py:72-99) with the K samples sharded across
    ranks: every rank runs its (sample, scene) pairs as one folded forward, the [K, S] per-scene error sums are
    all-reduced (10 KB for K = 20, S = 64) and the min over K / sum over scenes is taken on every rank.
    noise [K, S, *noise_dim] must be the same tensor on every rank (draw it from a common seed).  Every rank holds the
    whole minibatch (it is a few hundred pedestrians).  Same result as evaluate.evaluate_batch up to summation order."""
    from . import _lib
    from .ops import _f32, _ptr, _stream
    world = world if world is not None else (dist.get_world_size(group) if dist.is_initialized() else 1)
    rank = rank if rank is not None else (dist.get_rank(group) if dist.is_initialized() else 0)
    dev = obs_traj.device
    sched = get_schedule(seq_start_end, dev)
    sse = sched.host_sse
    S, K, T = sched.n_scenes, int(num_samples), pred_traj_gt.shape[0]
    plan = _sample_shard_plan(sched, K, world, rank, dev)
    sums = torch.zeros(2, K, S, dtype=torch.float32, device=dev)
    if plan is not None:
        idx, sse_l, sel_k, sel_s, pair_of_ped = plan
        z = noise.to(dev)[sel_k, sel_s]
        rel = generator(obs_traj.index_select(1, idx), obs_traj_rel.index_select(1, idx), sse_l,
                        obs_traj_g.index_select(1, idx), user_noise=z).contiguous()
        n_l = int(idx.numel())
        gt = _f32(pred_traj_gt.index_select(1, idx), 'pred_traj_gt')
        start = _f32(obs_traj[-1].index_select(0, idx), 'obs_traj')
        ade = torch.empty(n_l, 1, dtype=torch.float32, device=dev)
        fde = torch.empty_like(ade)
        L = _lib.lib()
        with torch.cuda.device(dev):
            _lib.check(L.sgx_displacement_errors(_ptr(rel), _ptr(start), _ptr(gt), T, n_l, _ptr(ade), _ptr(fde), 1, 0,
                                                 _stream(rel)), 'sgx_displacement_errors')
        sums.view(2, K * S).index_add_(1, pair_of_ped, torch.stack([ade.view(-1), fde.view(-1)], 0))
    if world > 1:
        dist.all_reduce(sums, op=dist.ReduceOp.SUM, group=group)
    best = sums.min(dim=1).values.sum(dim=1)
    return best[0], best[1]


def allreduce_gradients(module, group=None, ran_forward=True):
    """One all-reduce (sum) over every parameter gradient of `module`: ONE `cat` into a flat bucket, the collective,
    and the parameters' .grad re-pointed at views of the reduced bucket (no copy back).

    Parameters without a gradient (e.g. the passenger gcn_module of the 'gat' wiring) contribute zeros and get
    grad = None back, so an optimizer with weight decay / momentum skips them exactly as the reference's does.  The set
    of such parameters is a property of the wiring, identical on every rank that ran a forward; a rank with an EMPTY
    shard (`ran_forward=False`) contributes zeros for everything and reuses the set it saw on its last real step.
    Returns the number of bytes reduced (0 when there is nothing to do)."""
    params = [p for p in module.parameters() if p.requires_grad]
    if not params or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return 0
    if ran_forward:
        none_set = frozenset(i for i, p in enumerate(params) if p.grad is None)
        module._sgx_grad_none = none_set
    else:
        none_set = getattr(module, '_sgx_grad_none', frozenset())
    zeros = getattr(module, '_sgx_zero_grads', None)
    n_max = max(p.numel() for p in params)
    if zeros is None or zeros.numel() < n_max or zeros.device != params[0].device:
        zeros = torch.zeros(n_max, dtype=params[0].dtype, device=params[0].device)
        module._sgx_zero_grads = zeros
    pieces = [(p.grad.reshape(-1) if (ran_forward and p.grad is not None) else zeros[:p.numel()]) for p in params]
    flat = torch.cat(pieces)
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    off = 0
    for i, p in enumerate(params):
        n = p.numel()
        p.grad = None if i in none_set else flat[off:off + n].view_as(p)
        off += n
    return flat.numel() * flat.element_size()


def _global_count(n_local, device, group=None):
    """Fallback when the caller did not pass n_global: one extra (blocking) all-reduce per step."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return float(n_local)
    t = torch.tensor([float(n_local)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return float(t.item())


def _weight(n_local, n_global, device, group):
    total = float(n_global) if n_global else _global_count(n_local, device, group)
    return n_local / max(total, 1.0)


def variety_l2(l2_raw_per_sample, loss_mask, sched):
    """sum over scenes of min_k( sum_{peds of scene} l2[ped,k] ) / sum(loss_mask of scene)  -- the per-scene python
    loop of scripts/train.py:460-464 as two segmented sums."""
    seg = ped_scene_index(sched)
    k = l2_raw_per_sample.shape[1]
    per_scene = l2_raw_per_sample.new_zeros(sched.n_scenes, k).index_add_(0, seg, l2_raw_per_sample)
    denom = loss_mask.new_zeros(sched.n_scenes).index_add_(0, seg, loss_mask.sum(dim=1))
    return (per_scene.min(dim=1).values / denom).sum()


def discriminator_step(args, batch, generator, discriminator, optimizer_d, label_rng=None, group=None, n_global=None,
                       noise=None):
    """scripts/train.py:395-429 on this rank's shard; gradients are all-reduced before the optimizer step.
    n_global: pedestrians of the global minibatch (global_ped_count); without it one extra blocking all-reduce finds it."""
    (obs_traj, pred_traj_gt, obs_traj_rel, pred_traj_gt_rel, obs_traj_g, loss_mask, seq_start_end) = batch
    n_local = obs_traj.shape[1]
    w = _weight(n_local, n_global, obs_traj.device, group)
    if n_local == 0:                   # empty shard: no forward, zero gradients, every collective still joined
        if label_rng is not None:      # keep the label-smoothing stream aligned with the other ranks (gan_d_loss: 2 draws)
            label_rng.uniform(0.7, 1.2)
            label_rng.uniform(0, 0.3)
        optimizer_d.zero_grad()
        allreduce_gradients(discriminator, group, ran_forward=False)
        if getattr(args, 'clipping_threshold_d', 0) > 0:
            nn.utils.clip_grad_norm_(discriminator.parameters(), args.clipping_threshold_d)
        optimizer_d.step()
        return {'D_total_loss': 0.0}
    # The reference leaves the generator graph attached here (scripts/train.py:404-409) and back-propagates the D loss
    # into generator .grad buffers that optimizer_g.zero_grad() discards before they are ever used; running the
    # generator without autograd gives the same D gradients and takes the inference kernels.
    with torch.no_grad():    # noise: optional [rows, *noise_dim] for this rank's scenes (parity runs under sharding)
        fake_rel = generator(obs_traj, obs_traj_rel, seq_start_end, obs_traj_g, user_noise=noise)
        fake = relative_to_abs(fake_rel, obs_traj[-1])
    if _batch_independent(discriminator):
        # fake and real trajectories as ONE discriminator batch of 2 x scenes (same weights, independent scenes)
        traj = torch.cat([torch.cat([obs_traj, fake], 0), torch.cat([obs_traj, pred_traj_gt], 0)], 1)
        traj_rel = torch.cat([torch.cat([obs_traj_rel, fake_rel], 0), torch.cat([obs_traj_rel, pred_traj_gt_rel], 0)], 1)
        scores = discriminator(traj, traj_rel, tiled_schedule(seq_start_end, 2, obs_traj.device))
        s_fake, s_real = scores[:n_local], scores[n_local:]
    else:
        s_fake = discriminator(torch.cat([obs_traj, fake], 0), torch.cat([obs_traj_rel, fake_rel], 0), seq_start_end)
        s_real = discriminator(torch.cat([obs_traj, pred_traj_gt], 0), torch.cat([obs_traj_rel, pred_traj_gt_rel], 0),
                               seq_start_end)
    loss = gan_d_loss(s_real, s_fake, label_rng) * w
    optimizer_d.zero_grad()
    loss.backward()
    allreduce_gradients(discriminator, group)
    if getattr(args, 'clipping_threshold_d', 0) > 0:
        nn.utils.clip_grad_norm_(discriminator.parameters(), args.clipping_threshold_d)
    optimizer_d.step()
    # 0-dim tensors, not floats: converting here would synchronise the host with the device every step (the reference
    # only reads its losses every print_every iterations, scripts/train.py:316-330)
    return {'D_total_loss': loss.detach() / max(w, 1e-12)}


def _batch_independent(module):
    """True when every pedestrian / scene is processed independently of the rest of the batch, i.e. stacking batches is
    exact: no BatchNorm statistics and no active dropout."""
    for m in module.modules():
        if isinstance(m, nn.modules.batchnorm._BatchNorm):
            return False
        if isinstance(m, nn.Dropout) and m.p > 0 and m.training:
            return False
    return True


def _folded_samples(generator, obs_traj, obs_traj_rel, obs_traj_g, seq_start_end, k, noise=None):
    """k generator samples as one forward over k copies of the batch -> fake_rel [pred_len, k * batch, 2] (sample-major).
    noise: optional [k, rows, *noise_dim] (rows = scenes for 'global' mixing) instead of k get_noise draws."""
    from .models import get_noise
    n, s = obs_traj.shape[1], get_schedule(seq_start_end, obs_traj.device).n_scenes
    if noise is not None:
        noise = noise.reshape(-1, *noise.shape[2:]).to(obs_traj.device)
    elif generator.noise_dim:
        rows = s if generator.noise_mix_type == 'global' else n
        # the k draws of the sample loop, in order, on the CPU generator (seed-compatible with the reference) -- but ONE
        # host-to-device copy from pinned memory instead of k synchronous pageable ones
        noise = torch.cat([get_noise((rows,) + tuple(generator.noise_dim), generator.noise_type) for _ in range(k)], dim=0)
        if obs_traj.is_cuda:
            noise = noise.pin_memory().to(obs_traj.device, non_blocking=True)
    sse_k = tiled_schedule(seq_start_end, k, obs_traj.device)        # a SceneSchedule: no device read-back
    return generator(obs_traj.repeat(1, k, 1), obs_traj_rel.repeat(1, k, 1), sse_k, obs_traj_g.repeat(1, k, 1),
                     user_noise=noise)


def generator_step(args, batch, generator, discriminator, optimizer_g, label_rng=None, group=None, n_global=None,
                   noise=None):
    """scripts/train.py:432-484: best-of-K variety loss + adversarial term on the last sample."""
    (obs_traj, pred_traj_gt, obs_traj_rel, pred_traj_gt_rel, obs_traj_g, loss_mask, seq_start_end) = batch
    n_local = obs_traj.shape[1]
    w = _weight(n_local, n_global, obs_traj.device, group)
    if n_local == 0:                   # empty shard (see discriminator_step)
        if label_rng is not None:
            label_rng.uniform(0.7, 1.2)
        optimizer_g.zero_grad()
        allreduce_gradients(generator, group, ran_forward=False)
        if getattr(args, 'clipping_threshold_g', 0) > 0:
            nn.utils.clip_grad_norm_(generator.parameters(), args.clipping_threshold_g)
        optimizer_g.step()
        return {'G_total_loss': 0.0}
    sched = get_schedule(seq_start_end, obs_traj.device)
    mask = loss_mask[:, args.obs_len:]
    raws, raw_kn = [], None
    if getattr(args, 'fold_best_k', True) and args.best_k > 1 and _batch_independent(generator):
        # The best_k samples share weights and inputs and differ only in the noise: run them as ONE forward / backward
        # over best_k copies of the batch (SURVEY 8d cfg 4/5: "K folded into the batch dimension").  Same noise stream
        # as the loop (one get_noise draw per sample, in order), same loss terms, 1/best_k of the launches.
        fake_all = _folded_samples(generator, obs_traj, obs_traj_rel, obs_traj_g, seq_start_end, args.best_k, noise)
        fake_rel = fake_all[:, (args.best_k - 1) * n_local:]      # the last sample feeds the adversarial term (train.py:466)
        if args.l2_loss_weight > 0:
            # l2_loss(..., mode='raw') of all best_k samples in ONE expression (the same masked squared error, summed over
            # coordinates then time): the per-sample loop was 20 x (slice, sub, pow, mul, 2 sums) + as many autograd nodes,
            # ~300 launches and 3 ms of host time in a step that is launch-bound
            T = fake_all.shape[0]
            err = mask.unsqueeze(2).unsqueeze(0) * (pred_traj_gt_rel.permute(1, 0, 2).unsqueeze(0) -
                                                    fake_all.view(T, args.best_k, n_local, 2).permute(1, 2, 0, 3)) ** 2
            raw_kn = args.l2_loss_weight * err.sum(dim=3).sum(dim=2)                     # [best_k, n]
    else:
        for k in range(args.best_k):     # noise: optional [best_k, rows, *noise_dim] (parity runs under sharding)
            fake_rel = generator(obs_traj, obs_traj_rel, seq_start_end, obs_traj_g,
                                 user_noise=None if noise is None else noise[k])
            if args.l2_loss_weight > 0:
                raws.append(args.l2_loss_weight * l2_loss(fake_rel, pred_traj_gt_rel, mask, mode='raw'))
    loss = obs_traj.new_zeros(())
    losses = {}
    if args.l2_loss_weight > 0:
        l2 = variety_l2(raw_kn.t() if raw_kn is not None else torch.stack(raws, dim=1), mask, sched)
        losses['G_l2_loss_rel'] = l2.detach()
        loss = loss + l2
    fake = relative_to_abs(fake_rel, obs_traj[-1])
    s_fake = discriminator(torch.cat([obs_traj, fake], 0), torch.cat([obs_traj_rel, fake_rel], 0), seq_start_end)
    adv = gan_g_loss(s_fake, label_rng)
    losses['G_discriminator_loss'] = adv.detach()
    loss = loss + adv * w
    optimizer_g.zero_grad()
    loss.backward()
    allreduce_gradients(generator, group)
    if getattr(args, 'clipping_threshold_g', 0) > 0:
        nn.utils.clip_grad_norm_(generator.parameters(), args.clipping_threshold_g)
    optimizer_g.step()
    losses['G_total_loss'] = loss.detach()
    return losses


def make_label_rng(seed, step):
    """Identical on every rank: the reference draws label smoothing from Python's global RNG (sgan/losses.py:32,45)."""
    return random.Random(seed * 1000003 + step)
