"""CPU, world_size = 2 over gloo: the N > 1 host logic -- deterministic LPT sharding, flattened gradient all-reduce,
the BCE re-weighting rule, and the vectorised variety loss against the reference's per-scene loop."""
import os
import random
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import sse_from_sizes


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        from group_gan_gcn_gat_b200 import parallel
        from group_gan_gcn_gat_b200.losses import bce_loss
        rng = np.random.RandomState(3)
        sizes = list(rng.randint(2, 30, size=40))
        sse = sse_from_sizes(sizes)
        n = int(sse[-1, 1])
        torch.manual_seed(0)
        x = torch.randn(8, n, 2)
        scores_in = torch.randn(n, 5)
        local, local_sse, mine = parallel.shard_batch({'x': x, '0:s': scores_in}, sse, world, rank)
        # (1) shards are disjoint, cover everything, keep whole scenes, and are identical on every rank
        gathered = [None] * world
        dist.all_gather_object(gathered, mine.tolist())
        assert sorted(sum(gathered, [])) == list(range(len(sizes)))
        assert local['x'].shape[1] == int(local_sse[-1, 1]) == sum(sizes[s] for s in mine)
        first = int(sse[mine[0], 0])
        assert torch.equal(local['x'][:, :sizes[mine[0]]], x[:, first:first + sizes[mine[0]]])
        # (2) flattened gradient all-reduce with the BCE weighting == single-process gradient on the full batch
        torch.manual_seed(1)
        net = torch.nn.Sequential(torch.nn.Linear(5, 4), torch.nn.ReLU(), torch.nn.Linear(4, 1))
        target = torch.full((n,), 0.9)
        full_loss = bce_loss(net(scores_in).squeeze(1), target)
        full_grads = torch.autograd.grad(full_loss, list(net.parameters()))
        n_local = local['s'].shape[0]
        w = n_local / parallel._global_count(n_local, 'cpu')
        loss = bce_loss(net(local['s']).squeeze(1), target[:n_local]) * w
        net.zero_grad()
        loss.backward()
        nbytes = parallel.allreduce_gradients(net)
        assert nbytes == sum(p.numel() for p in net.parameters()) * 4
        for p, g in zip(net.parameters(), full_grads):
            assert torch.allclose(p.grad, g, atol=1e-6), (p.grad - g).abs().max()
        # (2b) no blocking count exchange needed: the global ped count comes from the global seq_start_end
        assert parallel.global_ped_count(sse.numpy()) == n
        assert abs(parallel._weight(n_local, n, 'cpu', None) - w) < 1e-12
        # (2c) fewer scenes than ranks: rank 1 gets an EMPTY shard, skips its forward, still joins the all-reduce and
        # ends with the same reduced gradients; a parameter nobody touched keeps grad = None on both ranks
        one = sse_from_sizes([5])
        loc1, sse1, mine1 = parallel.shard_batch({'x': x[:, :5]}, one, world, rank)
        assert (loc1['x'].shape[1], sse1.shape[0], len(mine1)) == ((5, 1, 1) if rank == 0 else (0, 0, 0))

        class Net(torch.nn.Module):
            def __init__(self):
                super().__init__()
                self.used = torch.nn.Linear(2, 3)
                self.passenger = torch.nn.Linear(2, 2)           # like gcn_module in the 'gat' wiring
        torch.manual_seed(5)
        net2 = Net()
        net2.zero_grad()
        if rank == 0:
            net2.used(loc1['x'][-1]).square().sum().backward()
            expect = [net2.used.weight.grad.clone(), net2.used.bias.grad.clone()]
            parallel.allreduce_gradients(net2)
        else:
            net2._sgx_grad_none = frozenset({2, 3})              # what this rank saw on its last non-empty step
            parallel.allreduce_gradients(net2, ran_forward=False)
            expect = None
        both = [None] * world
        dist.all_gather_object(both, [net2.used.weight.grad, net2.used.bias.grad])
        assert torch.equal(both[0][0], both[1][0]) and torch.equal(both[0][1], both[1][1])
        if rank == 0:
            assert torch.allclose(net2.used.weight.grad, expect[0]) and torch.allclose(net2.used.bias.grad, expect[1])
        assert net2.passenger.weight.grad is None and net2.passenger.bias.grad is None
        # (3) identically seeded label RNG
        vals = [None] * world
        dist.all_gather_object(vals, parallel.make_label_rng(7, 3).uniform(0.7, 1.2))
        assert vals[0] == vals[1]
        q.put((rank, 'ok'))
    except Exception as e:      # pragma: no cover
        q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


def test_world2_gloo_sharding_and_allreduce():
    import __graft_entry__ as ge
    ge.build()
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(r[1] == 'ok' for r in results), results


def test_variety_loss_matches_per_scene_loop():
    import __graft_entry__ as ge
    ge.build()
    from group_gan_gcn_gat_b200 import parallel
    from group_gan_gcn_gat_b200.schedule import SceneSchedule
    torch.manual_seed(2)
    sizes = [3, 1, 7, 2, 12]
    sse = sse_from_sizes(sizes)
    n, k = sum(sizes), 6
    raw = torch.rand(n, k)
    mask = (torch.rand(n, 12) > 0.1).float()
    ref = torch.zeros(())
    for s, e in sse.tolist():                       # scripts/train.py:460-464
        ref = ref + torch.min(raw[s:e].sum(dim=0)) / mask[s:e].sum()
    got = parallel.variety_l2(raw, mask, SceneSchedule(sse, 'cpu'))
    assert torch.allclose(got, ref, rtol=1e-6)


def test_losses_match_oracle_definitions():
    from group_gan_gcn_gat_b200 import losses
    from group_gan_gcn_gat_b200.utils import relative_to_abs
    from oracle import sgan_oracle as O
    torch.manual_seed(4)
    rel = torch.randn(12, 9, 2)
    start = torch.randn(9, 2)
    gt = torch.randn(12, 9, 2)
    ab = relative_to_abs(rel, start)
    assert torch.allclose(ab, O.relative_to_abs(rel, start), atol=1e-6)
    assert torch.allclose(losses.displacement_error(ab, gt, mode='raw'), O.displacement_error_raw(ab, gt), atol=1e-6)
    assert torch.allclose(losses.final_displacement_error(ab[-1], gt[-1], mode='raw'),
                          O.final_displacement_error_raw(ab[-1], gt[-1]), atol=1e-6)
    x, y = torch.randn(50) * 3, torch.rand(50)
    ref = torch.nn.functional.binary_cross_entropy_with_logits(x, y)
    assert torch.allclose(losses.bce_loss(x, y), ref, atol=1e-6)
    r = random.Random(1)
    a = losses.gan_d_loss(x, -x, r)
    assert torch.isfinite(a)


def test_sample_pair_partition_is_balanced_and_complete():
    """K-sample sharding: the (sample, scene) pairs of one 64-scene minibatch over 8 ranks -- every pair
    exactly once, N^2 cost balanced to within one large scene, identical for every caller."""
    import __graft_entry__ as ge
    ge.build()
    from group_gan_gcn_gat_b200 import parallel
    rng = np.random.RandomState(5)
    sizes = rng.randint(2, 40, size=64)
    sse = sse_from_sizes(list(sizes))
    K, world = 20, 8
    r = parallel.partition_samples(sse, K, world)
    assert r.shape == (K, 64) and r.min() == 0 and r.max() == world - 1
    assert np.array_equal(r, parallel.partition_samples(sse, K, world))
    cost = np.array([(np.tile(sizes.astype(np.int64) ** 2, (K, 1)) * (r == w)).sum() for w in range(world)])
    assert cost.sum() == K * (sizes.astype(np.int64) ** 2).sum()
    assert cost.max() - cost.min() <= int(sizes.max()) ** 2
    # scene sharding alone would leave the same batch unbalanced by far more
    r1 = parallel.partition_samples(sse, 1, world)
    cost1 = np.array([((sizes.astype(np.int64) ** 2) * (r1[0] == w)).sum() for w in range(world)])
    assert (cost.max() / cost.mean()) <= (cost1.max() / cost1.mean()) + 1e-9
