"""GPU parity: the CUDA path (through the C ABI / torch.library ops) against the golden vectors frozen from
the reference and against the CPU oracle on seeded inputs.  Tolerances (BASELINE.json north_star):
group structure bit-exact; pooled features 1e-5 relative (fp32); ADE/FDE 1e-4.
"""
import numpy as np
import pytest
import torch

from conftest import load_golden, sse_from_sizes, state_dict_of
from oracle import sgan_oracle as O

pytestmark = pytest.mark.gpu

DEV = 'cuda:0'
torch.backends.cudnn.allow_tf32 = False   # cuDNN LSTM (Encoder/Decoder) must not drop to TF32 for 1e-5 parity


def rel_err(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    assert a.shape == b.shape, (a.shape, b.shape)
    return ((a - b).abs().max() / max(1e-6, b.abs().max())).item()


def assert_close(a, b, tol, what='', floor=1e-6):
    """max |a-b| <= tol * max(|b|_max, floor).  `floor` guards gradients whose true value is ~0 (pure rounding
    noise in the reference, e.g. d/d(gat_inter.out_att.a) ~ 1e-7 next to O(10) sibling gradients)."""
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    assert a.shape == b.shape, (a.shape, b.shape)
    e = ((a - b).abs().max() / max(floor, b.abs().max())).item()
    assert e <= tol, '%s: relative error %.3e > %.1e' % (what, e, tol)


def grad_floor(g):
    """1e-2 of the largest parameter gradient of the fixture (cancellation noise scales with the terms, not the sum)."""
    return 1e-2 * max(float(v.abs().max()) for k, v in g.items() if k.startswith('grad.'))


@pytest.fixture(scope='module')
def sgx():
    import group_gan_gcn_gat_b200.modules as M
    import group_gan_gcn_gat_b200.models as MD
    import group_gan_gcn_gat_b200.ops as ops
    from group_gan_gcn_gat_b200.schedule import get_schedule
    return dict(M=M, MD=MD, ops=ops, get_schedule=get_schedule)


# ------------------------------------------------------------------ group structure (bit exact)
def test_group_structure_bit_exact_vs_golden(sgx):
    g = load_golden('groups')
    for case in sorted({k.split('.')[0] for k in g}):
        lab = g[case + '.labels'].reshape(-1)
        n = lab.numel()
        sched = sgx['get_schedule'](torch.tensor([[0, n]]), DEV)
        groups = sgx['ops'].group_ids(lab.to(DEV), sched.ped_start, sched.ped_end, sched.scene_start)
        M, A, R, Rn = sgx['ops'].group_dense(lab.to(DEV), groups, 0, n)
        assert torch.equal(M.cpu(), g[case + '.M']), case
        assert torch.equal(A.cpu(), g[case + '.A']), case          # fl(1/count) bit for bit
        assert torch.equal(R.cpu(), g[case + '.R']), case
        assert torch.equal(Rn.cpu(), g[case + '.Rn']), case
        assert int(groups[3][0]) == g[case + '.R'].shape[0]


def test_group_ids_vs_integer_oracle(sgx):
    rng = np.random.RandomState(7)
    sizes = list(rng.randint(1, 40, size=300)) + [257, 1]
    labs = np.concatenate([np.where(rng.rand(n) < 0.15, 0, rng.randint(1, max(2, n // 3 + 1), size=n)) for n in sizes])
    labs = labs.astype(np.float32)
    labs[5] = -0.0
    labs[100] = 2.5
    sse = sse_from_sizes(sizes)
    ref = O.group_ids_numpy(labs, sse)
    sched = sgx['get_schedule'](sse, DEV)
    leader, gsize, gid, ngrp = sgx['ops'].group_ids(torch.from_numpy(labs).to(DEV), sched.ped_start, sched.ped_end,
                                                    sched.scene_start)
    assert np.array_equal(leader.cpu().numpy(), ref['leader'])
    assert np.array_equal(gsize.cpu().numpy(), ref['group_size'])
    assert np.array_equal(gid.cpu().numpy(), ref['group_id'])
    assert np.array_equal(ngrp.cpu().numpy(), ref['n_group'])


# ------------------------------------------------------------------ generic gemm
def test_gemm_matches_torch(sgx):
    torch.manual_seed(0)
    for (m, n, k) in [(1, 1, 1), (70, 33, 5), (513, 40, 72), (24, 32, 10000), (3, 2, 70001)]:
        a = torch.randn(m, k, device=DEV)
        b = torch.randn(k, n, device=DEV)
        ref = (a.double() @ b.double()).float()
        assert_close(sgx['ops'].gemm(a, b), ref, 2e-6, 'gemm %s' % ((m, n, k),))
        assert_close(sgx['ops'].gemm(a.t().contiguous().t(), b.t().contiguous().t()), ref, 2e-6, 'gemm strided')


# ------------------------------------------------------------------ PoolHiddenNet
def _pool_module(sgx, sd, e_dim, h_dim, bott):
    m = sgx['M'].PoolHiddenNet(embedding_dim=e_dim, h_dim=h_dim, mlp_dim=64, bottleneck_dim=bott, batch_norm=False)
    m.load_state_dict(sd, strict=True)
    return m.to(DEV)


@pytest.mark.parametrize('name,dims', [('pool_g', (16, 32, 8)), ('pool_d', (16, 48, 48)), ('pool_g_big', (16, 32, 8))])
def test_pool_fwd_bwd_vs_golden(sgx, name, dims):
    g = load_golden(name)
    m = _pool_module(sgx, state_dict_of(g), *dims)
    h = g['h'].to(DEV).requires_grad_(True)
    pos = g['pos'].to(DEV).requires_grad_(True)
    out = m(h, g['seq_start_end'].to(DEV), pos)
    assert_close(out, g['out'], 1e-5, name + ' out')
    (out * g['upstream'].to(DEV)).sum().backward()
    assert_close(h.grad, g['grad_in.h'], 2e-5, name + ' dh')
    assert_close(pos.grad, g['grad_in.pos'], 2e-5, name + ' dpos')
    for k, p in m.named_parameters():
        assert_close(p.grad, g['grad.' + k], 2e-5, name + ' d' + k, floor=grad_floor(g))


@pytest.mark.parametrize('sizes,dims', [([1], (16, 32, 8)), ([2] * 300, (16, 32, 8)), ([129, 3, 200, 64, 65], (16, 32, 8)),
                                        ([40, 7, 130], (16, 48, 48)), ([5, 9], (8, 20, 24)), ([512], (16, 32, 8))])
def test_pool_vs_oracle_seeded(sgx, sizes, dims):
    e_dim, h_dim, bott = dims
    torch.manual_seed(len(sizes) * 31 + sizes[0])
    m = sgx['M'].PoolHiddenNet(embedding_dim=e_dim, h_dim=h_dim, mlp_dim=64, bottleneck_dim=bott, batch_norm=False)
    sse = sse_from_sizes(sizes)
    b = int(sse[-1, 1])
    h = torch.randn(1, b, h_dim)
    pos = torch.rand(b, 2) * 15
    ref = O.pool_hidden_net(h, sse, pos, m.state_dict())
    m = m.to(DEV)
    out = m(h.to(DEV), sse.to(DEV), pos.to(DEV))
    assert_close(out, ref, 1e-5, 'pool %s' % (sizes[:4],))


def test_pool_argmax_points_into_own_scene(sgx):
    g = load_golden('pool_g')
    sd = state_dict_of(g)
    sched = sgx['get_schedule'](g['seq_start_end'], DEV)
    out, arg = sgx['ops'].pool_fwd(g['h'].reshape(-1, 32).to(DEV), g['pos'].to(DEV), sched.ped_start, sched.ped_end,
                                   sched.pair_off, sched.tile_first, sched.n_pairs,
                                   *[sd[k].to(DEV) for k in ('spatial_embedding.weight', 'spatial_embedding.bias',
                                                             'mlp_pre_pool.0.weight', 'mlp_pre_pool.0.bias',
                                                             'mlp_pre_pool.2.weight', 'mlp_pre_pool.2.bias')], 0)
    ref_v, ref_i = O.pool_hidden_net_argmax(g['h'], g['seq_start_end'], g['pos'], sd)
    arg = arg.cpu()
    pos_mask = ref_v > 1e-6
    assert torch.equal(arg[pos_mask].long(), ref_i[pos_mask])
    for (s, e) in O.scene_bounds(g['seq_start_end']):
        assert ((arg[s:e] >= s) & (arg[s:e] < e)).all()


def test_pool_scene_independence_and_permutation_large(sgx):
    """Size-independent properties at dense-crowd size (N = 1024): block-diagonality and permutation equivariance."""
    torch.manual_seed(3)
    m = sgx['M'].PoolHiddenNet(embedding_dim=16, h_dim=32, mlp_dim=64, bottleneck_dim=8, batch_norm=False).to(DEV)
    n = 1024
    h = torch.randn(n + 37, 32, device=DEV)
    pos = torch.rand(n + 37, 2, device=DEV) * 15
    both = m(h, sse_from_sizes([n, 37]).to(DEV), pos)
    alone = m(h[:n], sse_from_sizes([n]).to(DEV), pos[:n])
    assert torch.equal(both[:n], alone)
    perm = torch.randperm(n, device=DEV)
    permuted = m(h[:n][perm], sse_from_sizes([n]).to(DEV), pos[:n][perm])
    assert_close(permuted, alone[perm], 1e-6, 'permutation equivariance')


def test_pool_rejects_bad_segments(sgx):
    m = sgx['M'].PoolHiddenNet(embedding_dim=16, h_dim=32, mlp_dim=64, bottleneck_dim=8, batch_norm=False).to(DEV)
    h = torch.randn(5, 32, device=DEV)
    pos = torch.rand(5, 2, device=DEV)
    with pytest.raises(ValueError):
        m(h, torch.tensor([[0, 2], [3, 5]]), pos)      # gap
    with pytest.raises(ValueError):
        m(h, torch.tensor([[0, 2], [2, 2], [2, 5]]), pos)  # empty scene
    with pytest.raises(ValueError):
        m(h, torch.tensor([[0, 4]]), pos)              # does not cover the batch


def test_graph_modules_reject_mismatched_feature_dims(sgx):
    """GATEncoder hard-codes GAT(40, ...) (sgan/models.py:242-244): a wiring with encoder_h_dim + bottleneck_dim != 40
    is a matmul size error in the reference and must be a ValueError here, not an out-of-bounds read in a kernel."""
    sse = sse_from_sizes([4, 3]).to(DEV)
    pos, labs = torch.rand(7, 2, device=DEV), torch.zeros(7, 1, device=DEV)
    gat = sgx['M'].GATEncoder(n_units=None, n_heads=1, dropout=0, alpha=0.2).to(DEV)
    for width in (32, 48):
        with pytest.raises(ValueError):
            gat(torch.randn(7, width, device=DEV), sse, pos, labs)
        with pytest.raises(ValueError):
            gat(torch.randn(7, width, device=DEV, requires_grad=True), sse, pos, labs)
    gcn = sgx['M'].GCNModule(input_dim=40, hidden_dim=72, out_dim=16, gcn_layers=2, final_dim=24).to(DEV)
    with pytest.raises(ValueError):
        gcn(torch.randn(7, 32, device=DEV), sse, pos, labs)


# ------------------------------------------------------------------ GATEncoder / GCNModule
@pytest.mark.parametrize('name', ['gat_encoder_h1', 'gat_encoder_h2'])
def test_gat_encoder_fwd_bwd_vs_golden(sgx, name):
    g = load_golden(name)
    m = sgx['M'].GATEncoder(n_units=None, n_heads=int(g['n_heads']), dropout=0, alpha=float(g['alpha']))
    m.load_state_dict(state_dict_of(g), strict=True)
    m = m.to(DEV)
    x = g['x'].to(DEV).requires_grad_(True)
    out = m(x, g['seq_start_end'].to(DEV), g['pos'].to(DEV), g['labels'].to(DEV))
    assert_close(out, g['out'], 1e-5, name + ' out')
    (out * g['upstream'].to(DEV)).sum().backward()
    assert_close(x.grad, g['grad_in.x'], 5e-5, name + ' dx')
    for k, p in m.named_parameters():
        assert_close(p.grad, g['grad.' + k], 5e-5, name + ' d' + k, floor=grad_floor(g))


@pytest.mark.parametrize('name,in_dim', [('gcn_module_40', 40), ('gcn_module_32', 32)])
def test_gcn_module_fwd_bwd_vs_golden(sgx, name, in_dim):
    g = load_golden(name)
    m = sgx['M'].GCNModule(input_dim=in_dim, hidden_dim=72, out_dim=16, gcn_layers=2, final_dim=24)
    m.load_state_dict(state_dict_of(g), strict=True)
    m = m.to(DEV)
    x = g['x'].to(DEV).requires_grad_(True)
    out = m(x, g['seq_start_end'].to(DEV), g['pos'].to(DEV), g['labels'].to(DEV))
    assert_close(out, g['out'], 1e-5, name + ' out')
    (out * g['upstream'].to(DEV)).sum().backward()
    assert_close(x.grad, g['grad_in.x'], 5e-5, name + ' dx')
    for k, p in m.named_parameters():
        assert_close(p.grad, g['grad.' + k], 5e-5, name + ' d' + k, floor=grad_floor(g))


@pytest.mark.parametrize('kind', ['gat', 'gcn'])
def test_graph_modules_vs_oracle_seeded(sgx, kind):
    """Ragged mix incl. singleton scenes, one big scene and one scene that is a single group."""
    rng = np.random.RandomState(11)
    torch.manual_seed(11)
    sizes = [1, 2, 57, 3, 300, 8, 1, 14]
    sse = sse_from_sizes(sizes)
    b = int(sse[-1, 1])
    labs = np.concatenate([np.where(rng.rand(n) < 0.13, 0, rng.randint(1, max(2, n // 3 + 1), size=n)) for n in sizes])
    labs[sse[5, 0]:sse[5, 1]] = 4          # the 8-ped scene is one group
    labs = torch.tensor(labs, dtype=torch.float32).view(-1, 1)
    x = torch.randn(b, 40)
    pos = torch.rand(b, 2)
    if kind == 'gat':
        m = sgx['M'].GATEncoder(n_units=None, n_heads=1, dropout=0, alpha=0.2)
        ref = O.gat_encoder(x, sse, pos, labs, m.state_dict(), '', 0.2, 1)
    else:
        m = sgx['M'].GCNModule(input_dim=40, hidden_dim=72, out_dim=16, gcn_layers=2, final_dim=24)
        with torch.no_grad():
            for p in m.parameters():
                if p.dim() == 2 and p.shape != (24, 32):
                    p.mul_(0.15)
        ref = O.gcn_module(x, sse, pos, labs, m.state_dict())
    out = m.to(DEV)(x.to(DEV), sse.to(DEV), pos.to(DEV), labs.to(DEV))
    assert_close(out, ref, 1e-5, kind + ' encoder')


# ------------------------------------------------------------------ generator / discriminator wiring
def _generator(sgx, g, wiring):
    MD = sgx['MD']
    pet = bool(int(g['pool_every_timestep']))
    gen = MD.TrajectoryGenerator(obs_len=8, pred_len=int(g['pred_len']), embedding_dim=16, encoder_h_dim=32,
                                 decoder_h_dim=32, mlp_dim=64, num_layers=1, noise_dim=(8,), noise_type='gaussian',
                                 noise_mix_type='global', pooling_type='pool_net', pool_every_timestep=pet, dropout=0,
                                 bottleneck_dim=8, batch_norm=False, n_heads=int(g['n_heads']), dropout1=0,
                                 alpha=float(g['alpha']), context_type=wiring)
    missing, unexpected = gen.load_state_dict(state_dict_of(g), strict=False)
    assert all(k.startswith('gcn_module') for k in unexpected), unexpected   # passenger of the reference class
    assert all(k.startswith(('mlp_decoder_context', 'gatencoder')) for k in missing), missing
    return gen.to(DEV).train()


def _set_pool_precision(gen, precision):
    for m in gen.modules():
        if isinstance(m, type(gen.pool_net)):
            m.precision = precision


# 'fp32' is what every module defaults to: the tensor-core kernel with fp16 hi/lo operand splits ('tc32') for the
# generator's pool_net dims; 'fp32-simt' pins the CUDA-core kernel.  Both must meet the ADE/FDE 1e-4 contract.
@pytest.mark.parametrize('precision', ['fp32', 'tc32', 'fp32-simt'])
@pytest.mark.parametrize('name', ['generator_gat_zara1', 'generator_p_eth', 'generator_gcn_zara1', 'generator_gat_pet'])
def test_generator_matches_reference(sgx, name, precision):
    g = load_golden(name)
    gen = _generator(sgx, g, str(g['wiring']))
    _set_pool_precision(gen, precision)
    obs, obs_rel, grp = g['obs_traj'].to(DEV), g['obs_traj_rel'].to(DEV), g['obs_traj_g'].to(DEV)
    sse = g['seq_start_end'].to(DEV)
    ades, fdes = [], []
    with torch.no_grad():
        for k in range(g['noise'].shape[0]):
            rel = gen(obs, obs_rel, sse, grp, user_noise=g['noise'][k].to(DEV))
            assert_close(rel, g['pred_rel'][k], 1e-4, name + ' pred_rel')   # 12 recurrent cuDNN steps; ADE/FDE below is the bar
            ab = O.relative_to_abs(rel.cpu(), g['obs_traj'][-1])
            ades.append(O.displacement_error_raw(ab, g['pred_traj_gt']))
            fdes.append(O.final_displacement_error_raw(ab[-1], g['pred_traj_gt'][-1]))
    n = obs.shape[1]
    ade = float(O.best_of_k(ades, g['seq_start_end'])) / (n * int(g['pred_len']))
    fde = float(O.best_of_k(fdes, g['seq_start_end'])) / n
    assert abs(ade - float(g['ade'])) < 1e-4 and abs(fde - float(g['fde'])) < 1e-4


@pytest.mark.parametrize('name', ['generator_gat_zara1', 'generator_p_eth', 'generator_gcn_zara1', 'generator_gat_pet'])
def test_generator_bf16_pooling_is_outside_the_ade_contract_but_bounded(sgx, name):
    """precision='bf16' (bf16 operands, 2e-2 pooled features) is NOT a contract mode for ADE/FDE: on the real
    checkpoints it moves ADE/FDE by ~1e-3.  It is kept as an explicit opt-in; this pins how far it is from the frozen
    reference numbers (bound 5e-3) so the documentation stays honest.  bench.py's headline runs 'fp32'."""
    g = load_golden(name)
    gen = _generator(sgx, g, str(g['wiring']))
    _set_pool_precision(gen, 'bf16')
    obs, obs_rel, grp = g['obs_traj'].to(DEV), g['obs_traj_rel'].to(DEV), g['obs_traj_g'].to(DEV)
    sse = g['seq_start_end'].to(DEV)
    ades, fdes = [], []
    with torch.no_grad():
        for k in range(g['noise'].shape[0]):
            rel = gen(obs, obs_rel, sse, grp, user_noise=g['noise'][k].to(DEV))
            ab = O.relative_to_abs(rel.cpu(), g['obs_traj'][-1])
            ades.append(O.displacement_error_raw(ab, g['pred_traj_gt']))
            fdes.append(O.final_displacement_error_raw(ab[-1], g['pred_traj_gt'][-1]))
    n = obs.shape[1]
    ade = float(O.best_of_k(ades, g['seq_start_end'])) / (n * int(g['pred_len']))
    fde = float(O.best_of_k(fdes, g['seq_start_end'])) / n
    print('bf16 pooling %s: |dADE| %.2e |dFDE| %.2e' % (name, abs(ade - float(g['ade'])), abs(fde - float(g['fde']))))
    assert abs(ade - float(g['ade'])) < 5e-3 and abs(fde - float(g['fde'])) < 1e-2


def test_discriminator_matches_reference(sgx):
    g = load_golden('discriminator_zara1')
    d = sgx['MD'].TrajectoryDiscriminator(obs_len=8, pred_len=12, embedding_dim=16, h_dim=48, mlp_dim=64, num_layers=1,
                                          dropout=0, batch_norm=False, d_type='global')
    d.load_state_dict(state_dict_of(g), strict=True)
    d = d.to(DEV)
    with torch.no_grad():
        s = d(g['traj'].to(DEV), g['traj_rel'].to(DEV), g['seq_start_end'].to(DEV))
    assert_close(s, g['scores'], 2e-5, 'D scores')


# ------------------------------------------------------------------ standalone dense layers
def test_dense_gat_layer_and_gat_vs_golden(sgx):
    M = sgx['M']
    g = load_golden('gat_layer_dense')
    layer = M.GraphAttentionLayer(12, 20, dropout=0, alpha=0.2, concat=True)
    layer.load_state_dict(state_dict_of(g), strict=True)
    layer = layer.to(DEV)
    x = g['x'].to(DEV).requires_grad_(True)
    out = layer(x, g['adj'].to(DEV))
    assert_close(out, g['out'], 1e-5, 'gat layer out')
    (out * g['upstream'].to(DEV)).sum().backward()
    assert_close(x.grad, g['grad_in.x'], 5e-5, 'gat layer dx')
    for k, p in layer.named_parameters():
        assert_close(p.grad, g['grad.' + k], 5e-5, 'gat layer d' + k, floor=grad_floor(g))
    g = load_golden('gat_dense')
    net = M.GAT(12, 20, 6, dropout=0, alpha=0.2, nheads=3)
    net.load_state_dict(state_dict_of(g), strict=True)
    net = net.to(DEV)
    x = g['x'].to(DEV).requires_grad_(True)
    out = net(x, g['adj'].to(DEV))
    assert_close(out, g['out'], 1e-5, 'gat out')
    (out * g['upstream'].to(DEV)).sum().backward()
    assert_close(x.grad, g['grad_in.x'], 5e-5, 'gat dx')
    for k, p in net.named_parameters():
        assert_close(p.grad, g['grad.' + k], 5e-5, 'gat d' + k, floor=grad_floor(g))


def test_dense_gat_fully_masked_row_is_uniform(sgx):
    """where(adj > 0, e, -9e15) then softmax: a row with no neighbour attends uniformly (models.py:202-204)."""
    torch.manual_seed(5)
    layer = sgx['M'].GraphAttentionLayer(6, 4, dropout=0, alpha=0.2, concat=False)
    x = torch.randn(5, 6)
    adj = torch.eye(5)
    adj[2] = 0
    ref = O.graph_attention_layer(x, adj, layer.W.detach(), layer.a.detach(), 0.2, False)
    out = layer.to(DEV)(x.to(DEV), adj.to(DEV))
    assert_close(out, ref, 1e-5, 'fully masked row')


def test_dense_gcn_vs_golden(sgx):
    g = load_golden('gcn_dense')
    net = sgx['M'].GCN(input_dim=12, hidden_dim=20, out_dim=6, gcn_layers=3)
    net.load_state_dict(state_dict_of(g), strict=True)
    net = net.to(DEV)
    x = g['x'].to(DEV).requires_grad_(True)
    out = net(g['adj'].to(DEV), x)
    assert_close(out, g['out'], 1e-5, 'gcn out')
    (out * g['upstream'].to(DEV)).sum().backward()
    assert_close(x.grad, g['grad_in.x'], 5e-5, 'gcn dx')
    for k, p in net.named_parameters():
        assert_close(p.grad, g['grad.' + k], 5e-5, 'gcn d' + k, floor=grad_floor(g))


# ------------------------------------------------------------------ fused LSTM recurrences
@pytest.mark.parametrize('h_dim,T', [(32, 8), (48, 20), (64, 3)])
def test_fused_encoder_matches_cudnn_and_oracle(sgx, h_dim, T):
    torch.manual_seed(h_dim)
    enc = sgx['MD'].Encoder(embedding_dim=16, h_dim=h_dim, mlp_dim=64, num_layers=1)
    x = torch.randn(T, 777, 2) * 0.4
    sd = {'e.' + k: v for k, v in enc.state_dict().items()}
    ref = O.encoder(x, sd, 'e.')
    enc = enc.to(DEV)
    with torch.no_grad():
        fused = enc(x.to(DEV))                      # inference -> sgx_lstm_encoder_fwd
    eager = enc(x.to(DEV).requires_grad_(True))     # autograd -> nn.LSTM
    assert fused.shape == eager.shape == ref.shape
    assert_close(fused, ref, 2e-5, 'fused encoder vs CPU oracle')   # ex2.approx-based sigmoid/tanh, T recurrent steps
    assert_close(fused, eager, 2e-5, 'fused encoder vs cuDNN')


@pytest.mark.parametrize('pet', [False, True])
def test_fused_decoder_matches_autograd_path(sgx, pet):
    torch.manual_seed(5)
    dec = sgx['MD'].Decoder(12, embedding_dim=16, h_dim=32, mlp_dim=64, num_layers=1, pool_every_timestep=pet,
                            bottleneck_dim=8, batch_norm=False).to(DEV)
    sizes = [3, 5, 2, 40, 1, 9]
    sse = sse_from_sizes(sizes).to(DEV)
    n = sum(sizes)
    last_pos = torch.rand(n, 2, device=DEV) * 10
    last_rel = torch.randn(n, 2, device=DEV) * 0.3
    h0 = torch.randn(1, n, 32, device=DEV)
    c0 = torch.zeros(1, n, 32, device=DEV)
    with torch.no_grad():
        fused, hf = dec(last_pos, last_rel, (h0, c0), sse)
    eager, he = dec(last_pos, last_rel, (h0.clone().requires_grad_(True), c0), sse)
    assert_close(fused, eager, 1e-5, 'fused decoder pred_rel')
    assert_close(hf, he, 1e-5, 'fused decoder final h')


# ------------------------------------------------------------------ training step (SURVEY cfg 3 / 5): fwd + bwd parity
@pytest.mark.parametrize('name', ['generator_gat_zara1', 'generator_gcn_zara1'])
def test_generator_training_gradients_match_oracle(sgx, name):
    """Best-of-K variety L2 + adversarial term through the pooled discriminator: every generator gradient on the GPU
    (sgx backward kernels + cuDNN LSTM) against CPU autograd through the oracle restatement."""
    from group_gan_gcn_gat_b200 import losses, parallel
    from group_gan_gcn_gat_b200.utils import relative_to_abs
    g = load_golden(name)
    gd = load_golden('discriminator_zara1')
    wiring = str(g['wiring'])
    gen = _generator(sgx, g, wiring)
    disc = sgx['MD'].TrajectoryDiscriminator(obs_len=8, pred_len=12, embedding_dim=16, h_dim=48, mlp_dim=64, num_layers=1,
                                             dropout=0, batch_norm=False, d_type='global')
    disc.load_state_dict(state_dict_of(gd), strict=True)
    disc = disc.to(DEV)
    obs, obs_rel, grp = g['obs_traj'], g['obs_traj_rel'], g['obs_traj_g']
    sse, gt = g['seq_start_end'], g['pred_traj_gt']
    gt_rel = torch.cat([(gt[0] - obs[-1]).unsqueeze(0), gt[1:] - gt[:-1]], 0)
    mask = torch.ones(obs.shape[1], 12)
    K = g['noise'].shape[0]

    def total_loss(gen_fn, disc_fn, variety_fn, dev):
        raws, rel = [], None
        for k in range(K):
            rel = gen_fn(g['noise'][k].to(dev))
            raws.append(losses.l2_loss(rel, gt_rel.to(dev), mask.to(dev), mode='raw'))
        l2 = variety_fn(torch.stack(raws, 1))
        fake = relative_to_abs(rel, obs[-1].to(dev))
        s = disc_fn(torch.cat([obs.to(dev), fake], 0), torch.cat([obs_rel.to(dev), rel], 0))
        return l2 + losses.bce_loss(s, torch.full_like(s, 0.9))

    # GPU: product path
    sched = sgx['get_schedule'](sse, DEV)
    loss_gpu = total_loss(lambda z: gen(obs.to(DEV), obs_rel.to(DEV), sse.to(DEV), grp.to(DEV), user_noise=z),
                          lambda t, tr: disc(t, tr, sse.to(DEV)), lambda r: parallel.variety_l2(r, mask.to(DEV), sched), DEV)
    gen.zero_grad()
    loss_gpu.backward()
    # CPU: oracle with autograd
    sd = {k: v.clone().requires_grad_(True) for k, v in state_dict_of(g).items()}
    sdd = state_dict_of(gd)
    cfg = dict(pred_len=12, wiring=wiring, pooling=True, pool_every_timestep=False, alpha=float(g['alpha']),
               n_heads=int(g['n_heads']))

    def variety_cpu(r):
        tot = 0.0
        for s, e in sse.tolist():
            tot = tot + torch.min(r[s:e].sum(0)) / mask[s:e].sum()
        return tot
    loss_cpu = total_loss(lambda z: O.generator_forward(obs, obs_rel, sse, grp, sd, cfg, z),
                          lambda t, tr: O.discriminator_forward(t, tr, sse, sdd), variety_cpu, 'cpu')
    loss_cpu.backward()
    assert abs(float(loss_gpu) - float(loss_cpu)) < 1e-4 * max(1.0, abs(float(loss_cpu)))
    ref_grads = {k: v.grad for k, v in sd.items() if v.grad is not None}
    floor = 1e-2 * max(float(v.abs().max()) for v in ref_grads.values())
    checked = 0
    # The max over neighbours is discontinuous in its sub-gradient: with the trained weights some (i, channel)
    # have two neighbours tied to ~1e-6, and the 6e-6 difference between cuDNN's and the CPU's LSTM output is enough
    # to flip the argmax, which moves O(1) of that event's gradient between pairs (measured: the same pooling op,
    # fed identical inputs, matches the oracle to 5e-7 -- see test_pool_fwd_bwd_vs_golden).  Parameters upstream
    # of that choice (first pooling layer, its embedding, the encoder) are therefore checked loosely here.
    flip_sensitive = ('pool_net.spatial_embedding', 'pool_net.mlp_pre_pool.0', 'encoder.')
    for k, p in gen.named_parameters():
        if k in ref_grads and p.grad is not None:
            tol = 8e-2 if k.startswith(flip_sensitive) else 1e-3
            assert_close(p.grad, ref_grads[k], tol, name + ' d' + k, floor=floor)
            checked += 1
    assert checked >= 20


def test_evaluate_batch_matches_reference_metrics(sgx):
    """scripts/evaluate_model.py:72-99 on one batch: best-of-K ADE/FDE (vectorised) against the frozen reference numbers."""
    from group_gan_gcn_gat_b200.evaluate import evaluate_batch
    g = load_golden('generator_gat_zara1')
    gen = _generator(sgx, g, 'gat')
    t = lambda k: g[k].to(DEV)
    n = g['obs_traj'].shape[1]
    for fold in (True, False):          # samples folded into one forward (the small-batch default) / the sample loop
        ade, fde = evaluate_batch(gen, t('obs_traj'), t('obs_traj_rel'), t('seq_start_end'), t('obs_traj_g'),
                                  t('pred_traj_gt'), num_samples=g['noise'].shape[0], noise=t('noise'), fold_samples=fold)
        assert abs(float(ade) / (n * 12) - float(g['ade'])) < 1e-4, fold
        assert abs(float(fde) / n - float(g['fde'])) < 1e-4, fold
    ade_h, fde_h = evaluate_batch(gen, t('obs_traj'), t('obs_traj_rel'), t('seq_start_end'), t('obs_traj_g'),
                                  t('pred_traj_gt'), num_samples=g['noise'].shape[0], noise=t('noise'), hoist_context=True)
    # hoisting gives bit-identical predictions (row f2); the scene sums are float atomics, so compare to rounding
    assert abs(float(ade_h) - float(ade)) <= 2e-6 * abs(float(ade)) and abs(float(fde_h) - float(fde)) <= 2e-6 * abs(float(fde))


# ------------------------------------------------------------------ tensor-core LSTM (3-way bf16 splits, fp32-level accuracy)
def _with_option(name, val, fn):
    """run fn with a library switch flipped (sgx_set_option), e.g. the CUDA-core recurrence kernels for every batch"""
    from group_gan_gcn_gat_b200 import _lib
    _lib.set_option(name, val)
    try:
        return fn()
    finally:
        _lib.set_option(name, 1)


def test_lstm_tensor_core_encoder_matches_cuda_core_and_oracle(sgx):
    torch.manual_seed(21)
    enc = sgx['MD'].Encoder(embedding_dim=16, h_dim=32, mlp_dim=64, num_layers=1)
    n = 128 * 77 + 5                                   # >= 8192 -> tcgen05 path; ragged last tile
    x = torch.randn(8, n, 2) * 0.4
    sd = {'e.' + k: v for k, v in enc.state_dict().items()}
    ref = O.encoder(x[:, :600], sd, 'e.')
    enc = enc.to(DEV)
    with torch.no_grad():
        tc = enc(x.to(DEV))
        cc = _with_option('lstm_tc', 0, lambda: enc(x.to(DEV)))
    assert_close(tc[:, :600], ref, 2e-5, 'tensor-core encoder vs CPU oracle')
    assert_close(tc, cc, 2e-5, 'tensor-core encoder vs CUDA-core kernel')


def test_lstm_tensor_core_saturated_gates(sgx):
    """The tensor-core kernel evaluates the cell update as a rational function of exponentials with clamped arguments:
    drive the gates deep into saturation (pre-activations of a few hundred, cell states of tens) and compare with
    torch's own LSTM on the CPU.  No NaN / Inf, same values."""
    torch.manual_seed(23)
    enc = sgx['MD'].Encoder(embedding_dim=16, h_dim=32, mlp_dim=64, num_layers=1)
    with torch.no_grad():
        for p in enc.encoder.parameters():
            p.mul_(12.0)
        enc.spatial_embedding.weight.mul_(6.0)
    n = 128 * 70 + 17
    x = torch.randn(8, n, 2) * 25.0
    x[:, ::7] = 0.0                                      # and some exactly-zero tracks
    import copy
    with torch.no_grad():
        e64 = copy.deepcopy(enc).double()
        emb = e64.spatial_embedding(x.double().reshape(-1, 2)).view(8, n, 16)
        ref = e64.encoder(emb)[1][0].float()             # float64 nn.LSTM on the CPU: the true value
        emb32 = enc.spatial_embedding(x.reshape(-1, 2)).view(8, n, 16)
        cpu32 = enc.encoder(emb32)[1][0]                 # torch's own fp32 result, for scale
        pre = emb.abs().max()
    assert float(pre) > 100
    enc = enc.to(DEV)
    with torch.no_grad():
        tc = enc(x.to(DEV))
        cc = _with_option('lstm_tc', 0, lambda: enc(x.to(DEV)))
    assert bool(torch.isfinite(tc).all()) and bool(torch.isfinite(cc).all())
    # Unsaturated gates are differences of terms of magnitude ~300 here, so ANY fp32 evaluation carries ~1e-4 of rounding
    # (torch's fp32 CPU LSTM included): the bar is "as accurate as fp32 gets", not the well-conditioned 2e-5.
    e_tc, e_cc = float((tc.cpu() - ref).abs().max()), float((cc.cpu() - ref).abs().max())
    e_cpu32 = float((cpu32 - ref).abs().max())
    assert e_tc < 3e-4 and e_cc < 3e-4 and e_tc < 4 * max(e_cpu32, 2e-5), (e_tc, e_cc, e_cpu32)


def test_lstm_tensor_core_decoder_matches_cuda_core(sgx):
    g = load_golden('generator_gat_zara1')
    gen = _generator(sgx, g, 'gat')
    torch.manual_seed(22)
    n_scenes = 3000
    sizes = [int(v) for v in torch.randint(2, 7, (n_scenes,))]
    sse = sse_from_sizes(sizes).to(DEV)
    n = sum(sizes)
    assert n >= 8192
    ctx = torch.randn(n, 24, device=DEV)
    obs = torch.rand(8, n, 2, device=DEV) * 10
    obs_rel = torch.randn(8, n, 2, device=DEV) * 0.3
    z = torch.randn(n_scenes, 8, device=DEV)
    with torch.no_grad():
        tc = gen.decode(ctx, obs, obs_rel, sse, user_noise=z)
        cc = _with_option('lstm_tc', 0, lambda: gen.decode(ctx, obs, obs_rel, sse, user_noise=z))
    assert tc.shape == (12, n, 2)
    assert_close(tc, cc, 3e-5, 'tensor-core decoder vs CUDA-core kernel')
    # and against the autograd (cuDNN) path on a slice of whole scenes
    k = 200
    m = int(sse[k - 1, 1])
    eager = gen.decode(ctx[:m].clone().requires_grad_(True), obs[:, :m], obs_rel[:, :m], sse[:k].clone(), user_noise=z[:k])
    assert_close(tc[:, :m], eager, 3e-5, 'tensor-core decoder vs cuDNN path')


def test_gcn_fused_single_launch_matches_three_kernel_path(sgx):
    g = load_golden('gcn_module_40')
    m = sgx['M'].GCNModule(input_dim=40, hidden_dim=72, out_dim=16, gcn_layers=2, final_dim=24)
    m.load_state_dict(state_dict_of(g), strict=True)
    m = m.to(DEV)
    args = (g['x'].to(DEV), g['seq_start_end'].to(DEV), g['pos'].to(DEV), g['labels'].to(DEV))
    import os
    with torch.no_grad():
        os.environ['SGX_GCN_FUSED'] = '0'                                   # host-side switch of GCNModule._chunks
        try:
            ref = m(*args)
        finally:
            os.environ.pop('SGX_GCN_FUSED', None)
        fused = m(*args)                                                    # default: tensor-core single launch
    assert_close(fused, ref, 5e-6, 'fused GCN (mma) vs three-kernel path')
    assert_close(fused, g['out'], 1e-5, 'fused GCN vs golden')


# ------------------------------------------------------------------ edge cases of the ragged layout
@pytest.mark.parametrize('sizes', [[32], [32, 32, 1], [33], [31, 2, 32, 1, 1, 30], [1] * 70, [64], [57, 3, 40, 24, 1],
                                   [64, 64, 1, 63], [65], [20, 70, 5]])
def test_gat_encoder_chunk_boundaries(sgx, sizes):
    """Scenes of exactly 32 peds fill a warp-chunk of the fused kernel; 33..64 take the two-slots-per-lane kernel
    (chunks of <= 64); 65 falls back to the general path; many singleton scenes share a chunk.  All paths must
    agree with the oracle."""
    rng = np.random.RandomState(sum(sizes))
    torch.manual_seed(sum(sizes))
    sse = sse_from_sizes(sizes)
    n = sum(sizes)
    labs = torch.tensor(np.where(rng.rand(n) < 0.2, 0, rng.randint(1, 5, size=n)), dtype=torch.float32).view(-1, 1)
    x, pos = torch.randn(n, 40), torch.rand(n, 2)
    m = sgx['M'].GATEncoder(n_units=None, n_heads=1, dropout=0, alpha=0.2)
    ref = O.gat_encoder(x, sse, pos, labs, m.state_dict(), '', 0.2, 1)
    with torch.no_grad():
        out = m.to(DEV)(x.to(DEV), sse.to(DEV), pos.to(DEV), labs.to(DEV))
    assert_close(out, ref, 1e-5, 'gat %s' % sizes[:3])


@pytest.mark.parametrize('scale', [1e-3, 30.0])
def test_graph_kernels_tensor_core_splits_at_extreme_magnitudes(sgx, scale):
    """The single-launch GAT / GCN forwards run their linear maps as 3xTF32 tensor-core GEMMs (operands split hi + lo):
    tiny and large activations (softmax saturation, log-softmax of large logits, ReLU sums of large terms) must still
    match the fp32 oracle to the 1e-5 contract relative to the output scale."""
    rng = np.random.RandomState(9)
    torch.manual_seed(9)
    sizes = [3, 32, 7, 1, 19, 12, 28, 4]
    sse = sse_from_sizes(sizes)
    n = sum(sizes)
    labs = torch.tensor(np.where(rng.rand(n) < 0.2, 0, rng.randint(1, 4, size=n)), dtype=torch.float32).view(-1, 1)
    x, pos = torch.randn(n, 40) * scale, torch.rand(n, 2)
    gat = sgx['M'].GATEncoder(n_units=None, n_heads=1, dropout=0, alpha=0.2)
    gcn = sgx['M'].GCNModule(input_dim=40, hidden_dim=72, out_dim=16, gcn_layers=2, final_dim=24)
    with torch.no_grad():
        for p in gcn.parameters():
            if p.dim() == 2 and p.shape[0] != 24:
                p.mul_(0.15)
    ref_gat = O.gat_encoder(x, sse, pos, labs, gat.state_dict(), '', 0.2, 1)
    ref_gcn = O.gcn_module(x, sse, pos, labs, gcn.state_dict(), '')
    with torch.no_grad():
        out_gat = gat.to(DEV)(x.to(DEV), sse.to(DEV), pos.to(DEV), labs.to(DEV))
        out_gcn = gcn.to(DEV)(x.to(DEV), sse.to(DEV), pos.to(DEV), labs.to(DEV))
    assert bool(torch.isfinite(out_gat).all()) and bool(torch.isfinite(out_gcn).all())
    assert_close(out_gat, ref_gat, 1e-5, 'gat at scale %g' % scale)
    assert_close(out_gcn, ref_gcn, 1e-5, 'gcn at scale %g' % scale)


@pytest.mark.parametrize('sizes,in_dim,final', [([32], 40, 24), ([32, 32, 1], 32, 24), ([33], 40, 24),
                                                ([31, 2, 32, 1, 1, 30], 40, 32), ([1] * 70, 32, 32)])
def test_gcn_module_chunk_boundaries(sgx, sizes, in_dim, final):
    """the same layout edge cases through the single-launch tensor-core GCN kernel (and its fallback for N = 33), for
    every built (input_dim, final_dim) instance"""
    rng = np.random.RandomState(sum(sizes) + in_dim)
    torch.manual_seed(sum(sizes) + final)
    sse = sse_from_sizes(sizes)
    n = sum(sizes)
    labs = torch.tensor(np.where(rng.rand(n) < 0.2, 0, rng.randint(1, 5, size=n)), dtype=torch.float32).view(-1, 1)
    x, pos = torch.randn(n, in_dim), torch.rand(n, 2)
    m = sgx['M'].GCNModule(input_dim=in_dim, hidden_dim=72, out_dim=16, gcn_layers=2, final_dim=final)
    with torch.no_grad():
        for p in m.parameters():
            if p.dim() == 2 and p.shape[0] != final:
                p.mul_(0.15)                                   # plain randn GCN weights: keep activations O(1)
    ref = O.gcn_module(x, sse, pos, labs, m.state_dict(), '')
    with torch.no_grad():
        out = m.to(DEV)(x.to(DEV), sse.to(DEV), pos.to(DEV), labs.to(DEV))
    assert_close(out, ref, 1e-5, 'gcn %s' % sizes[:3])


@pytest.mark.parametrize('sizes', [[8] * 2, [16] * 8, [2] * 32, [11, 3], [128], [1]])
def test_pool_tile_boundaries_both_precisions(sgx, sizes):
    """Pair counts that are exact multiples of the 128-pair tile, one pair short / over, and a single pair."""
    torch.manual_seed(len(sizes))
    m = sgx['M'].PoolHiddenNet(embedding_dim=16, h_dim=32, mlp_dim=64, bottleneck_dim=8, batch_norm=False)
    sse = sse_from_sizes(sizes)
    n = sum(sizes)
    h, pos = torch.randn(n, 32), torch.rand(n, 2) * 15
    ref = O.pool_hidden_net(h, sse, pos, m.state_dict())
    m = m.to(DEV)
    with torch.no_grad():
        for precision in ('fp32', 'fp32-simt', 'tc32'):
            m.precision = precision
            assert_close(m(h.to(DEV), sse.to(DEV), pos.to(DEV)), ref, 1e-5, '%s %s' % (precision, sizes[:2]))
        m.precision = 'bf16'
        assert_close(m(h.to(DEV), sse.to(DEV), pos.to(DEV)), ref, 2e-2, 'bf16 %s' % sizes[:2])


def test_generator_every_timestep_pooling_bf16_close_to_fp32(sgx):
    """cfg 4 wiring (decoder pools at every step): the bf16 tensor-core pooling stays within 2e-2 of the fp32 path
    after 12 recurrent steps on a dense scene."""
    g = load_golden('generator_gat_pet')
    gen = _generator(sgx, g, 'gat')
    torch.manual_seed(3)
    n = 96
    sse = sse_from_sizes([n]).to(DEV)
    obs = torch.rand(8, n, 2, device=DEV) * 10
    obs_rel = torch.randn(8, n, 2, device=DEV) * 0.2
    grp = torch.randint(0, 6, (8, n, 1), device=DEV).float()
    z = torch.randn(1, 8, device=DEV)
    with torch.no_grad():
        a = gen(obs, obs_rel, sse, grp, user_noise=z)
        gen.pool_net.precision = gen.decoder.pool_net.precision = 'bf16'
        b = gen(obs, obs_rel, sse, grp, user_noise=z)
    assert_close(b, a, 2e-2, 'per-step pooling bf16 vs fp32')


def test_generator_block_diagonal_over_scenes_at_bench_size(sgx):
    """Size-independent property at BASELINE's full size (65 536 zara1-shaped scenes, 245 k peds): the generator is
    block-diagonal over scenes, so the forward of the whole batch equals the concatenated forwards of its two halves,
    in both pooling precisions, and the LPT shards of `parallel.shard_batch` reassemble to the same predictions."""
    import sys
    import os
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench
    from group_gan_gcn_gat_b200 import parallel
    data = bench.synth_batch(1 << 16, 4321)
    g = load_golden('generator_gat_zara1')
    gen = _generator(sgx, g, 'gat')
    sse = data['seq_start_end']
    S = sse.shape[0]
    n = int(sse[-1, 1])
    z = torch.randn(S, 8, generator=torch.Generator().manual_seed(5))
    dev = lambda t: t.to(DEV)

    def forward(scenes_lo, scenes_hi):
        p0, p1 = int(sse[scenes_lo, 0]), int(sse[scenes_hi - 1, 1])
        part = sse[scenes_lo:scenes_hi] - p0
        with torch.no_grad():
            return gen(dev(data['obs_traj'][:, p0:p1]), dev(data['obs_traj_rel'][:, p0:p1]), dev(part),
                       dev(data['obs_traj_g'][:, p0:p1]), user_noise=dev(z[scenes_lo:scenes_hi]))

    for precision, tol in (('bf16', 1e-6), ('fp32', 1e-6), ('fp32-simt', 1e-6)):
        gen.pool_net.precision = precision
        whole = forward(0, S)
        assert whole.shape == (12, n, 2) and bool(torch.isfinite(whole).all())
        halves = torch.cat([forward(0, S // 2 + 7), forward(S // 2 + 7, S)], dim=1)
        assert_close(halves, whole, tol, 'halves vs whole (%s)' % precision)
    # LPT shards over 3 ranks: every pedestrian appears exactly once and keeps its prediction
    gen.pool_net.precision = 'bf16'
    whole = forward(0, S)
    seen = torch.zeros(n, dtype=torch.bool)
    for rank in range(3):
        loc, sse_r, mine = parallel.shard_batch({'obs_traj': data['obs_traj'], 'obs_traj_rel': data['obs_traj_rel'],
                                                 'obs_traj_g': data['obs_traj_g']}, sse, 3, rank)
        mine = torch.as_tensor(mine)
        with torch.no_grad():
            out = gen(dev(loc['obs_traj']), dev(loc['obs_traj_rel']), dev(sse_r), dev(loc['obs_traj_g']), user_noise=dev(z[mine]))
        idx = torch.cat([torch.arange(int(sse[s, 0]), int(sse[s, 1])) for s in mine.tolist()])
        assert not seen[idx].any()
        seen[idx] = True
        assert_close(out, whole[:, idx.to(DEV)], 1e-6, 'LPT shard %d' % rank)
    assert bool(seen.all())


# ------------------------------------------------------------------ fused context MLP (make_mlp at models.py:898, 165-166, 990)
@pytest.mark.parametrize('dims', [(32, 8, 64, 24), (32, 8, 64, 32), (32, 0, 64, 24), (48, 0, 64, 1), (40, 8, 64, 32), (24, 8, 64, 1)])
@pytest.mark.parametrize('batch', [1, 31, 32, 33, 1000, 70001])
def test_fused_mlp_matches_sequential(sgx, dims, batch):
    """ops.mlp2 == make_mlp([...])(cat([xa, xb], 1)) (Linear, ReLU, Linear, ReLU; sgan/models.py:7-20) to 1e-5, for every
    built (in, mid, out), batches around the 32-row warp chunk, with and without the folded concatenation."""
    da, db, hid, out = dims
    torch.manual_seed(da + db + out + batch)
    seq = sgx['M'].make_mlp([da + db, hid, out], activation='relu', batch_norm=False, dropout=0).to(DEV)
    xa = torch.randn(batch, da, device=DEV)
    xb = torch.randn(batch, db, device=DEV) if db else None
    with torch.no_grad():
        ref = seq(xa if xb is None else torch.cat([xa, xb], 1))
        got = sgx['ops'].mlp2(seq, xa, xb)
    assert got is not None and got.shape == ref.shape
    assert_close(got, ref.double(), 1e-5, 'mlp2 %s' % (dims,))


def test_fused_mlp_declines_what_it_cannot_do(sgx):
    M, ops = sgx['M'], sgx['ops']
    x = torch.randn(10, 40, device=DEV)
    assert ops.mlp2(M.make_mlp([40, 64, 24], batch_norm=True).to(DEV), x) is None            # BatchNorm inside
    assert ops.mlp2(M.make_mlp([40, 128, 24], batch_norm=False).to(DEV), x) is None          # mid width not built
    assert ops.mlp2(M.make_mlp([40, 64, 24], activation='leakyrelu', batch_norm=False).to(DEV), x) is None
    seq = M.make_mlp([40, 64, 24], batch_norm=False).to(DEV)
    assert ops.mlp2(seq, x.clone().requires_grad_(True)) is None                             # autograd -> torch path
    with torch.no_grad():
        assert ops.mlp2(seq, x) is not None


# ------------------------------------------------------------------ single-launch GATEncoder backward
@pytest.mark.parametrize('sizes', [[32], [32, 32, 1], [31, 2, 32, 1, 1, 30], [1] * 70, [3, 2, 7, 13, 4, 1], [2] * 500, [14, 9, 1, 27]])
def test_gat_encoder_fused_backward_vs_oracle_autograd(sgx, sizes):
    """Scenes <= 32 peds take gat_fused_bwd_kernel (forward recomputed per warp chunk, 3xTF32 warp GEMMs for dX and dW,
    per-CTA gradient blocks): every gradient against torch autograd through the CPU oracle of sgan/models.py:254-294,
    and against the general multi-pass backward."""
    rng = np.random.RandomState(sum(sizes) + len(sizes))
    torch.manual_seed(sum(sizes))
    sse = sse_from_sizes(sizes)
    n = sum(sizes)
    labs = torch.tensor(np.where(rng.rand(n) < 0.2, 0, rng.randint(1, 5, size=n)), dtype=torch.float32).view(-1, 1)
    x, pos, up = torch.randn(n, 40), torch.rand(n, 2), torch.randn(n, 24)
    m = sgx['M'].GATEncoder(n_units=None, n_heads=1, dropout=0, alpha=0.2)
    sd = {k: v.clone().requires_grad_(True) for k, v in m.state_dict().items()}
    xr = x.clone().requires_grad_(True)
    (O.gat_encoder(xr, sse, pos, labs, sd, '', 0.2, 1) * up).sum().backward()
    m = m.to(DEV)
    xg = x.to(DEV).requires_grad_(True)
    out = m(xg, sse.to(DEV), pos.to(DEV), labs.to(DEV))
    (out * up.to(DEV)).sum().backward()
    floor = 1e-2 * max(float(v.grad.abs().max()) for v in sd.values())
    assert_close(xg.grad, xr.grad, 5e-5, "dx %s" % sizes[:3], floor=floor)
    for k, p in m.named_parameters():
        # the attention vectors' gradients are sums of terms that cancel to ~1e-3 of their size (the reference's own
        # fp32 result moves by this much under a different summation order): 1e-4 of the largest gradient for those
        assert_close(p.grad, sd[k].grad, 1e-4 if k.endswith('.a') else 5e-5, 'd%s %s' % (k, sizes[:3]), floor=floor)
    # the general path on the same inputs (n_chunks = 0)
    from group_gan_gcn_gat_b200.schedule import get_schedule
    sched = get_schedule(sse, DEV)
    groups = sgx['ops'].group_ids(labs.to(DEV).reshape(-1), sched.ped_start, sched.ped_end, sched.scene_start)
    Wi, ai, Wio, aio = m.gat_intra.stacked()
    We, ae, Weo, aeo = m.gat_inter.stacked()
    ps = [t.detach() for t in (Wi, ai, Wio, aio, We, ae, Weo, aeo, m.out_embedding.weight, m.out_embedding.bias)]
    chunk_scene, n_chunks = sched.chunks(32)
    args = (xg.detach(), up.to(DEV), groups[0], groups[1], sched.ped_start, sched.ped_end, sched.n_scenes, *ps, 0.2,
            sched.scene_start, chunk_scene)
    fused = sgx['ops'].gat_encoder_bwd(*args, n_chunks, 32)
    general = sgx['ops'].gat_encoder_bwd(*args, 0, 32)
    assert n_chunks > 0
    for a, b in zip(fused, general):
        assert_close(a, b, 1e-4, 'fused vs general backward', floor=floor)


# ------------------------------------------------------------------ dense crowds (BASELINE.json configs[3])
@pytest.mark.parametrize('n', [256, 1024])
def test_dense_crowd_modules_vs_reference_goldens(sgx, n):
    """ONE scene of 256 / 1024 pedestrians through the unmodified reference modules (PoolHiddenNet materialises
    N^2 x 512, GATEncoder [N,N,144]) frozen by oracle/make_golden_r2.py: pooled features (tensor-core and CUDA-core
    kernels), GATEncoder (warp-per-row scene kernels) and GCNModule outputs within 1e-5."""
    g = load_golden('pool_g_%d' % n)
    m = _pool_module(sgx, state_dict_of(g), 16, 32, 8)
    with torch.no_grad():
        for precision in ('fp32', 'fp32-simt'):
            m.precision = precision
            assert_close(m(g['h'].to(DEV), g['seq_start_end'].to(DEV), g['pos'].to(DEV)), g['out'], 1e-5,
                         'pool N=%d %s' % (n, precision))
        m.precision = 'bf16'
        assert_close(m(g['h'].to(DEV), g['seq_start_end'].to(DEV), g['pos'].to(DEV)), g['out'], 2e-2, 'pool bf16 N=%d' % n)
    g = load_golden('gat_encoder_%d' % n)
    gat = sgx['M'].GATEncoder(n_units=None, n_heads=1, dropout=0, alpha=0.2)
    gat.load_state_dict(state_dict_of(g), strict=True)
    with torch.no_grad():
        out = gat.to(DEV)(g['x'].to(DEV), g['seq_start_end'].to(DEV), g['pos'].to(DEV), g['labels'].to(DEV))
    assert_close(out, g['out'], 1e-5, 'GATEncoder N=%d' % n)
    g = load_golden('gcn_module_%d' % n)
    gcn = sgx['M'].GCNModule(input_dim=40, hidden_dim=72, out_dim=16, gcn_layers=2, final_dim=24)
    gcn.load_state_dict(state_dict_of(g), strict=True)
    with torch.no_grad():
        out = gcn.to(DEV)(g['x'].to(DEV), g['seq_start_end'].to(DEV), g['pos'].to(DEV), g['labels'].to(DEV))
    assert_close(out, g['out'], 1e-5, 'GCNModule N=%d' % n)


@pytest.mark.parametrize('sizes', [[65], [200, 70, 129], [300, 3, 2000], [97] * 9])
def test_dense_crowd_gat_scene_kernels_vs_thread_kernels(sgx, sizes):
    """The warp-per-row scene kernels (max scene 65 .. 2048) against the thread-per-node general path on the same
    inputs, forward and backward (the backward's forward recompute takes the scene kernels too), mixed scene sizes."""
    rng = np.random.RandomState(sum(sizes))
    torch.manual_seed(sum(sizes))
    from group_gan_gcn_gat_b200.schedule import get_schedule
    sse = sse_from_sizes(sizes)
    n = sum(sizes)
    labs = torch.tensor(np.where(rng.rand(n) < 0.15, 0, rng.randint(1, max(2, max(sizes) // 3), size=n)),
                        dtype=torch.float32).view(-1, 1).to(DEV)
    x, up = torch.randn(n, 40, device=DEV), torch.randn(n, 24, device=DEV)
    m = sgx['M'].GATEncoder(n_units=None, n_heads=1, dropout=0, alpha=0.2).to(DEV)
    sched = get_schedule(sse, DEV)
    groups = sgx['ops'].group_ids(labs.reshape(-1), sched.ped_start, sched.ped_end, sched.scene_start)
    Wi, ai, Wio, aio = m.gat_intra.stacked()
    We, ae, Weo, aeo = m.gat_inter.stacked()
    ps = [t.detach() for t in (Wi, ai, Wio, aio, We, ae, Weo, aeo, m.out_embedding.weight, m.out_embedding.bias)]
    empty = sched.scene_start[:0]
    fwd = lambda max_scene: sgx['ops'].gat_encoder_fwd(x, groups[0], groups[1], sched.ped_start, sched.ped_end, sched.n_scenes,
                                                       *ps, 0.2, sched.scene_start, empty, 0, 32, max_scene)
    bwd = lambda max_scene: sgx['ops'].gat_encoder_bwd(x, up, groups[0], groups[1], sched.ped_start, sched.ped_end,
                                                       sched.n_scenes, *ps, 0.2, sched.scene_start, empty, 0, 32, max_scene)
    with torch.no_grad():
        assert_close(fwd(int(sched.max_n)), fwd(0), 1e-5, 'dense fwd %s' % sizes[:3])
        ref = bwd(0)
        floor = 1e-2 * max(float(b.abs().max()) for b in ref)      # d(gat_inter.out_att.a) is pure cancellation noise (~1e-6)
        for i, (a, b) in enumerate(zip(bwd(int(sched.max_n)), ref)):
            # outputs 2, 4, 6, 8 are the attention vectors' gradients: sums that cancel to ~1e-6 of their terms, so the
            # summation order of the forward recompute shows (both paths are within 1e-5 of autograd in absolute terms)
            assert_close(a, b, 1e-3 if i in (2, 4, 6, 8) else 5e-5, 'dense bwd %d %s' % (i, sizes[:3]), floor=floor)


# ------------------------------------------------------------------ single-launch GCNModule backward
@pytest.mark.parametrize('sizes,in_dim,final', [([32], 40, 24), ([32, 32, 1], 32, 24), ([31, 2, 32, 1, 1, 30], 40, 32),
                                                ([1] * 70, 32, 32), ([3, 2, 7, 13, 4, 1], 40, 24), ([2] * 500, 40, 24)])
def test_gcn_module_fused_backward_vs_oracle_autograd(sgx, sizes, in_dim, final):
    """Scenes <= 32 peds take gcn_fused_bwd_kernel (forward recomputed per warp chunk, collapsed group / scene rows, 3xTF32
    warp GEMMs): every gradient against torch autograd through the CPU oracle of sgan/models.py:628-712 and against the
    general multi-kernel backward."""
    rng = np.random.RandomState(sum(sizes) + in_dim)
    torch.manual_seed(sum(sizes) + final)
    sse = sse_from_sizes(sizes)
    n = sum(sizes)
    labs = torch.tensor(np.where(rng.rand(n) < 0.2, 0, rng.randint(1, 5, size=n)), dtype=torch.float32).view(-1, 1)
    x, pos, up = torch.randn(n, in_dim), torch.rand(n, 2), torch.randn(n, final)
    m = sgx['M'].GCNModule(input_dim=in_dim, hidden_dim=72, out_dim=16, gcn_layers=2, final_dim=final)
    with torch.no_grad():
        for p in m.parameters():
            if p.dim() == 2 and p.shape[0] != final:
                p.mul_(0.15)
    sd = {k: v.clone().requires_grad_(True) for k, v in m.state_dict().items()}
    xr = x.clone().requires_grad_(True)
    (O.gcn_module(xr, sse, pos, labs, sd, '') * up).sum().backward()
    m = m.to(DEV)
    xg = x.to(DEV).requires_grad_(True)
    (m(xg, sse.to(DEV), pos.to(DEV), labs.to(DEV)) * up.to(DEV)).sum().backward()
    floor = 1e-2 * max(float(v.grad.abs().max()) for v in sd.values())
    assert_close(xg.grad, xr.grad, 5e-5, 'dx %s' % sizes[:3], floor=floor)
    for k, p in m.named_parameters():
        assert_close(p.grad, sd[k].grad, 5e-5, 'd%s %s' % (k, sizes[:3]), floor=floor)
    from group_gan_gcn_gat_b200.schedule import get_schedule
    sched = get_schedule(sse, DEV)
    groups = sgx['ops'].group_ids(labs.to(DEV).reshape(-1), sched.ped_start, sched.ped_end, sched.scene_start)
    ps = [t.detach() for t in (m.gcn_intra.W[0], m.gcn_intra.W[1], m.gcn_inter.W[0], m.gcn_inter.W[1],
                               m.out_embedding.weight, m.out_embedding.bias)]
    chunk_scene, n_chunks = sched.chunks(32)
    args = (xg.detach(), up.to(DEV), groups[0], groups[1], sched.ped_start, sched.ped_end, sched.scene_start, groups[3], *ps,
            chunk_scene)
    assert n_chunks > 0
    for a, b in zip(sgx['ops'].gcn_module_bwd(*args, n_chunks), sgx['ops'].gcn_module_bwd(*args, 0)):
        assert_close(a, b, 5e-5, 'fused vs general GCN backward', floor=floor)
