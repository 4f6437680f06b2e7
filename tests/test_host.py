"""CPU: host-side logic -- the C ABI surface, the scene schedule, the module mirror's state_dict layout,
and the no-fallback guarantee.  No kernel is launched here."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from conftest import ROOT, load_golden, sse_from_sizes, state_dict_of


@pytest.fixture(scope='module')
def lib():
    import __graft_entry__ as ge
    ge.build()
    from group_gan_gcn_gat_b200 import _lib
    return _lib


def _declared_functions():
    """(name, n_params) for every prototype in include/sgx.h."""
    text = open(os.path.join(ROOT, 'include', 'sgx.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    protos = re.findall(r'\b(?:int|int64_t|long long|const char\*)\s+(sgx_\w+)\s*\(([^;]*?)\)\s*;', text, flags=re.S)
    out = []
    for name, params in protos:
        params = params.strip()
        n = 0 if params in ('', 'void') else params.count(',') + 1
        out.append((name, n))
    return out


def test_c_abi_exports_every_declared_symbol(lib):
    decl = _declared_functions()
    assert len(decl) >= 18
    handle = ctypes.CDLL(lib.LIB_PATH)
    for name, n_params in decl:
        assert hasattr(handle, name), 'libsgx_b200.so does not export %s' % name
        assert name in lib.SIGNATURES, 'python binding missing for %s' % name
        assert len(lib.SIGNATURES[name][1]) == n_params, '%s: header has %d params, binding %d' % (
            name, n_params, len(lib.SIGNATURES[name][1]))
    assert set(lib.SIGNATURES) == {n for n, _ in decl}
    assert lib.lib().sgx_version() >= 100


def test_library_carries_sm100a_code(lib):
    import shutil
    import subprocess
    exe = shutil.which('cuobjdump') or '/usr/local/cuda/bin/cuobjdump'
    if not os.path.isfile(exe):
        pytest.skip('cuobjdump not available')
    out = subprocess.run([exe, '-lelf', lib.LIB_PATH], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True).stdout
    assert 'sm_100a' in out


def test_schedule_matches_numpy(lib):
    from group_gan_gcn_gat_b200.schedule import SceneSchedule
    rng = np.random.RandomState(0)
    sizes = [1] + list(rng.randint(1, 70, size=500)) + [300, 2]
    sse = sse_from_sizes(sizes)
    s = SceneSchedule(sse, 'cpu')
    n = np.asarray(sizes)
    assert s.batch == n.sum() and s.max_n == n.max() and s.n_pairs == (n * n).sum() and s.n_scenes == len(sizes)
    starts = np.concatenate([[0], np.cumsum(n)])
    assert np.array_equal(s.scene_start.numpy(), starts)
    assert np.array_equal(s.ped_start.numpy(), np.repeat(starts[:-1], n))
    assert np.array_equal(s.ped_end.numpy(), np.repeat(starts[1:], n))
    per_ped = np.repeat(n, n)
    off = np.concatenate([[0], np.cumsum(per_ped)])
    assert np.array_equal(s.pair_off.numpy(), off)
    # tile t starts at pair 128 t, owned by the last ped whose offset is <= 128 t
    t = np.arange(s.n_tiles) * 128
    assert np.array_equal(s.tile_first.numpy(), np.searchsorted(off, t, side='right') - 1)


def test_schedule_rejects_what_the_reference_cannot_slice(lib):
    from group_gan_gcn_gat_b200.schedule import SceneSchedule
    for bad in ([[0, 2], [3, 5]], [[1, 3]], [[0, 2], [2, 2]], [[0, 3], [2, 5]]):
        with pytest.raises(ValueError):
            SceneSchedule(torch.tensor(bad), 'cpu')


def test_schedule_cache_follows_tensor_identity_and_version(lib):
    from group_gan_gcn_gat_b200.schedule import get_schedule
    a = torch.tensor([[0, 2], [2, 5]])
    s1 = get_schedule(a, 'cpu')
    assert get_schedule(a, 'cpu') is s1
    a[1, 1] = 6
    s2 = get_schedule(a, 'cpu')
    assert s2 is not s1 and s2.batch == 6
    assert get_schedule(a.clone(), 'cpu') is not s2


def test_lpt_partition_balances_n_squared(lib):
    from group_gan_gcn_gat_b200.schedule import SceneSchedule
    rng = np.random.RandomState(1)
    sizes = list(rng.randint(2, 60, size=512))
    s = SceneSchedule(sse_from_sizes(sizes), 'cpu')
    for world in (1, 2, 8):
        rank, cost = s.partition(world)
        assert rank.min() == 0 and rank.max() == world - 1
        c = np.asarray(sizes) ** 2
        assert np.array_equal(cost, np.bincount(rank, weights=c, minlength=world).astype(np.int64))
        assert cost.max() - cost.min() <= c.max()          # LPT bound
    r1, _ = s.partition(8)
    r2, _ = s.partition(8)
    assert np.array_equal(r1, r2)                          # deterministic: every rank computes the same split


def test_state_dict_layout_matches_reference_checkpoints(lib):
    """strict=True loading of the reference's own parameter names."""
    import group_gan_gcn_gat_b200.models as MD
    import group_gan_gcn_gat_b200.modules as M
    g = load_golden('generator_gat_zara1')
    gen = MD.TrajectoryGenerator(obs_len=8, pred_len=12, embedding_dim=16, encoder_h_dim=32, decoder_h_dim=32, mlp_dim=64,
                                 noise_dim=(8,), noise_mix_type='global', pooling_type='pool_net',
                                 pool_every_timestep=False, bottleneck_dim=8, batch_norm=False, n_heads=1)
    gen.load_state_dict(state_dict_of(g), strict=True)
    assert sum(p.numel() for p in gen.parameters()) == 56810
    g = load_golden('discriminator_zara1')
    d = MD.TrajectoryDiscriminator(obs_len=8, pred_len=12, embedding_dim=16, h_dim=48, mlp_dim=64, batch_norm=False,
                                   d_type='global')
    d.load_state_dict(state_dict_of(g), strict=True)
    assert sum(p.numel() for p in d.parameters()) == 73873
    g = load_golden('generator_gat_pet')
    gen = MD.TrajectoryGenerator(obs_len=8, pred_len=12, embedding_dim=16, encoder_h_dim=32, decoder_h_dim=32, mlp_dim=64,
                                 noise_dim=(8,), noise_mix_type='global', pooling_type='pool_net',
                                 pool_every_timestep=True, bottleneck_dim=8, batch_norm=False, n_heads=1)
    gen.load_state_dict(state_dict_of(g), strict=True)                # decoder.pool_net.*, decoder.mlp.*
    for name, cls, kw in [('gat_encoder_h2', M.GATEncoder, dict(n_units=None, n_heads=2, dropout=0, alpha=0.2)),
                          ('gcn_module_32', M.GCNModule, dict(input_dim=32, hidden_dim=72, out_dim=16, gcn_layers=2, final_dim=24)),
                          ('pool_d', M.PoolHiddenNet, dict(embedding_dim=16, h_dim=48, mlp_dim=64, bottleneck_dim=48, batch_norm=False)),
                          ('gat_dense', M.GAT, dict(nfeat=12, nhid=20, nclass=6, dropout=0, alpha=0.2, nheads=3)),
                          ('gcn_dense', M.GCN, dict(input_dim=12, hidden_dim=20, out_dim=6, gcn_layers=3))]:
        cls(**kw).load_state_dict(state_dict_of(load_golden(name)), strict=True)


def test_no_cpu_fallback(lib):
    """The product path must fail loudly without CUDA -- it never routes through the oracle or torch eager."""
    import group_gan_gcn_gat_b200.modules as M
    m = M.PoolHiddenNet(embedding_dim=16, h_dim=32, mlp_dim=64, bottleneck_dim=8, batch_norm=False)
    with pytest.raises(RuntimeError):
        m(torch.randn(1, 5, 32), torch.tensor([[0, 5]]), torch.rand(5, 2))
    enc = M.GCNModule()
    with pytest.raises((RuntimeError, TypeError)):
        enc(torch.randn(5, 40), torch.tensor([[0, 5]]), torch.rand(5, 2), torch.zeros(5, 1))
    src = ''
    pkg = os.path.join(ROOT, 'group_gan_gcn_gat_b200')
    for fn in os.listdir(pkg):
        if fn.endswith('.py'):
            src += open(os.path.join(pkg, fn)).read()
    assert 'oracle' not in src.replace('the oracle', '')       # the package never imports the checker


def test_unsupported_options_raise(lib):
    import group_gan_gcn_gat_b200.modules as M
    m = M.PoolHiddenNet(embedding_dim=16, h_dim=32, mlp_dim=64, bottleneck_dim=8, batch_norm=True)
    with pytest.raises(NotImplementedError):
        m(torch.randn(1, 5, 32), torch.tensor([[0, 5]]), torch.rand(5, 2))
    import group_gan_gcn_gat_b200.models as MD
    with pytest.raises(ValueError):
        MD.get_noise((2, 2), 'laplace')


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the CPU port on the host cores) needs no GPU: one JSON line with the contract keys."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, 'bench.py'), '--impl', 'reference', '--steps', '1', '--warmup', '1',
                          '--ref-scenes', '32'],
                         capture_output=True, text=True, timeout=600, cwd=root)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    for key in ('metric', 'value', 'unit', 'n_gpus', 'steps', 'warmup', 'ms_per_step', 'higher_is_better', 'scaling',
                'vs_baseline', 'dtype', 'data', 'config', 'impl', 'cpu_baseline', 'e2e'):
        assert key in line, key
    assert line['impl'] == 'reference' and line['metric'] == 'predicted_trajectories_per_sec' and line['value'] > 0
    assert line['cpu_baseline']['kind'] == 'port' and line['cpu_baseline']['cores'] >= 1
    assert line['e2e'] == {'value': line['value'], 'unit': line['unit'], 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}
    assert 'workload' in line['config']


def test_schedule_build_matches_the_separate_passes():
    """sgx_schedule_build (one pass, what SceneSchedule uses) == sgx_schedule_fill + per-ped scene index +
    sgx_schedule_chunks, on ragged layouts with and without a scene above the chunk capacity."""
    from group_gan_gcn_gat_b200 import _lib
    L = _lib.lib()
    rng = np.random.RandomState(3)
    # (>= 8192 scenes: the per-pedestrian pass runs on several host threads)
    for sizes in ([1], [32, 1, 31, 2], list(rng.randint(1, 15, size=500)), [5, 33, 2], [64] * 3 + [1],
                  list(rng.randint(1, 15, size=20001)), [1] * 8192, list(rng.randint(1, 40, size=9000))):
        st = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
        sse = np.ascontiguousarray(np.stack([st[:-1], st[1:]], 1))
        S, B = len(sizes), int(st[-1])
        stats = np.zeros(8, np.int64)
        assert L.sgx_schedule_stats(sse.ctypes.data, S, stats.ctypes.data) == 0
        T = max(int(stats[3]), 1)
        mk = lambda: (np.full(S + 1, -7, np.int32), np.full(B, -7, np.int32), np.full(B, -7, np.int32),
                      np.full(B + 1, -7, np.int64), np.full(T, -7, np.int32))
        a, b = mk(), mk()
        assert L.sgx_schedule_fill(sse.ctypes.data, S, *[x.ctypes.data for x in a]) == 0
        ped_scene, chunks, n = np.full(B, -7, np.int32), np.full(S + 1, -7, np.int32), np.zeros(1, np.int64)
        assert L.sgx_schedule_build(sse.ctypes.data, S, *[x.ctypes.data for x in b], ped_scene.ctypes.data, 32,
                                    chunks.ctypes.data, n.ctypes.data) == 0
        for x, y in zip(a, b):
            assert np.array_equal(x, y)
        assert np.array_equal(ped_scene, np.repeat(np.arange(S, dtype=np.int32), sizes))
        if max(sizes) <= 32:
            ref, m = np.full(S + 1, -7, np.int32), np.zeros(1, np.int64)
            assert L.sgx_schedule_chunks(sse.ctypes.data, S, 32, ref.ctypes.data, m.ctypes.data) == 0
            assert n[0] == m[0] and np.array_equal(chunks[:n[0] + 1], ref[:m[0] + 1])
            fill = np.add.reduceat(np.asarray(sizes), chunks[:n[0]])
            assert fill.max() <= 32
        else:
            assert n[0] == 0 and (chunks == -7).all()          # a scene exceeds the capacity: untouched


def test_batch_independence_guard_of_the_folded_steps():
    """parallel._batch_independent: folding samples / stacking fake + real is only exact without batch statistics"""
    import torch.nn as nn
    from group_gan_gcn_gat_b200.parallel import _batch_independent
    assert _batch_independent(nn.Sequential(nn.Linear(4, 4), nn.ReLU(), nn.Dropout(0.0)))
    assert not _batch_independent(nn.Sequential(nn.Linear(4, 4), nn.BatchNorm1d(4)))
    drop = nn.Sequential(nn.Linear(4, 4), nn.Dropout(0.5))
    assert not _batch_independent(drop.train())
    assert _batch_independent(drop.eval())


def test_tiled_schedule_is_the_schedule_of_k_copies():
    """schedule.tiled_schedule(sse, k): k copies of the batch side by side, built from the host copy of the base schedule
    (what the K-folded forwards and the stacked fake + real discriminator batch run on) == the schedule of the explicitly
    repeated seq_start_end; cached on the base schedule."""
    from group_gan_gcn_gat_b200.schedule import SceneSchedule, get_schedule, tiled_schedule
    rng = np.random.RandomState(5)
    sizes = list(rng.randint(1, 20, size=37))
    st = np.concatenate([[0], np.cumsum(sizes)])
    sse = torch.from_numpy(np.stack([st[:-1], st[1:]], 1).astype(np.int64))
    n = int(st[-1])
    for k in (1, 2, 20):
        t = tiled_schedule(sse, k, 'cpu')
        rep = sse.repeat(k, 1) + (torch.arange(k) * n).repeat_interleave(len(sizes)).unsqueeze(1)
        ref = SceneSchedule(rep, 'cpu')
        assert (t.n_scenes, t.batch, t.n_pairs, t.max_n) == (ref.n_scenes, ref.batch, ref.n_pairs, ref.max_n)
        for name in ('scene_start', 'ped_start', 'ped_end', 'pair_off', 'tile_first'):
            assert torch.equal(getattr(t, name), getattr(ref, name)), name
        assert tiled_schedule(sse, k, 'cpu') is t                                  # cached on the base schedule
        assert get_schedule(t, 'cpu') is t                                         # accepted wherever seq_start_end is


def test_ready_is_a_no_op_for_tensors_that_were_not_staged():
    """utils.ready (the wait on a staged host-batch copy) must leave ordinary tensors alone: the generator calls it on
    every input"""
    from group_gan_gcn_gat_b200.utils import ready, ready_last
    t = torch.arange(6.).view(2, 3)
    assert ready(t) is t and ready_last(t) is t and not hasattr(t, '_sgx_ready') and not hasattr(t, '_sgx_ready_last')
