"""CPU: the oracle restatement (oracle/sgan_oracle.py) against the golden vectors frozen from the reference."""
import numpy as np
import pytest
import torch

from conftest import load_golden, state_dict_of
from oracle import sgan_oracle as O


def _close(a, b, tol=2e-6):
    assert a.shape == b.shape
    err = (a - b).abs().max().item()
    scale = max(1.0, b.abs().max().item())
    assert err <= tol * scale, (err, scale)


@pytest.mark.parametrize('name', ['pool_g', 'pool_d', 'pool_g_big'])
def test_pool_matches_reference(name):
    g = load_golden(name)
    sd = state_dict_of(g)
    h = g['h'].clone().requires_grad_(True)
    pos = g['pos'].clone().requires_grad_(True)
    params = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    out = O.pool_hidden_net(h, g['seq_start_end'], pos, params)
    _close(out.detach(), g['out'])
    (out * g['upstream']).sum().backward()
    _close(h.grad, g['grad_in.h'], 1e-5)
    _close(pos.grad, g['grad_in.pos'], 1e-5)
    for k, p in params.items():
        _close(p.grad, g['grad.' + k], 1e-5)
    v, idx = O.pool_hidden_net_argmax(g['h'], g['seq_start_end'], g['pos'], sd)
    _close(v, g['out'])
    for (s, e) in O.scene_bounds(g['seq_start_end']):
        assert ((idx[s:e] >= s) & (idx[s:e] < e)).all()


@pytest.mark.parametrize('name', ['gat_encoder_h1', 'gat_encoder_h2'])
def test_gat_encoder_matches_reference(name):
    g = load_golden(name)
    sd = {k: v.clone().requires_grad_(True) for k, v in state_dict_of(g).items()}
    x = g['x'].clone().requires_grad_(True)
    out = O.gat_encoder(x, g['seq_start_end'], g['pos'], g['labels'], sd, '', float(g['alpha']), int(g['n_heads']))
    _close(out.detach(), g['out'])
    (out * g['upstream']).sum().backward()
    _close(x.grad, g['grad_in.x'], 1e-5)
    for k, p in sd.items():
        _close(p.grad, g['grad.' + k], 1e-5)


@pytest.mark.parametrize('name', ['gcn_module_40', 'gcn_module_32'])
def test_gcn_module_matches_reference(name):
    g = load_golden(name)
    sd = {k: v.clone().requires_grad_(True) for k, v in state_dict_of(g).items()}
    x = g['x'].clone().requires_grad_(True)
    out = O.gcn_module(x, g['seq_start_end'], g['pos'], g['labels'], sd)
    _close(out.detach(), g['out'])
    (out * g['upstream']).sum().backward()
    _close(x.grad, g['grad_in.x'], 1e-5)
    for k, p in sd.items():
        _close(p.grad, g['grad.' + k], 1e-5)


def test_dense_layers_match_reference():
    g = load_golden('gat_layer_dense')
    sd = state_dict_of(g)
    _close(O.graph_attention_layer(g['x'], g['adj'], sd['W'], sd['a'], 0.2, True), g['out'])
    g = load_golden('gat_dense')
    _close(O.gat(g['x'], g['adj'], state_dict_of(g), '', 0.2, 3), g['out'])
    g = load_golden('gcn_dense')
    _close(O.gcn(g['adj'], g['x'], state_dict_of(g), '', 3), g['out'])


def test_group_structure_bit_exact():
    g = load_golden('groups')
    cases = sorted({k.split('.')[0] for k in g})
    assert 'notebook' in cases and 'gcnpy128' in cases
    for c in cases:
        lab = g[c + '.labels']
        m = O.group_mask(lab)
        assert torch.equal(m, g[c + '.M'])
        assert torch.equal(O.row_normalize(m), g[c + '.A'])
        r = O.group_rows(m)
        assert torch.equal(r, g[c + '.R'])
        assert torch.equal(O.row_normalize(r), g[c + '.Rn'])
        # integer oracle agrees with the dense matrices
        n = lab.shape[0]
        ids = O.group_ids_numpy(lab.numpy(), [[0, n]])
        assert ids['n_group'][0] == r.shape[0]
        for i in range(n):
            row = r[ids['group_id'][i]]
            assert bool(row[i])
            assert int(row.sum()) == ids['group_size'][i]
            assert int(torch.nonzero(row)[0]) == ids['leader'][i]


def test_known_answers():
    """Untitled.ipynb:546-556 (R matrix) and sgan/GCN.py:128 (8 groups ordered by min member)."""
    g = load_golden('groups')
    assert g['notebook.R'].int().tolist() == [[1, 1, 0, 0], [0, 0, 1, 0], [0, 0, 0, 1]]
    ids = O.group_ids_numpy(g['gcnpy128.labels'].numpy(), [[0, 10]])
    assert ids['n_group'][0] == 8
    assert sorted(set(ids['leader'].tolist())) == [0, 1, 2, 4, 6, 7, 8, 9]


@pytest.mark.parametrize('name', ['generator_gat_zara1', 'generator_p_eth', 'generator_gcn_zara1', 'generator_gat_pet'])
def test_generator_matches_reference(name):
    g = load_golden(name)
    sd = state_dict_of(g)
    cfg = dict(pred_len=int(g['pred_len']), wiring=str(g['wiring']), pooling=True,
               pool_every_timestep=bool(int(g['pool_every_timestep'])), alpha=float(g['alpha']),
               n_heads=int(g['n_heads']))
    ades, fdes = [], []
    for k in range(g['noise'].shape[0]):
        rel = O.generator_forward(g['obs_traj'], g['obs_traj_rel'], g['seq_start_end'], g['obs_traj_g'], sd, cfg,
                                  g['noise'][k])
        _close(rel, g['pred_rel'][k], 1e-5)
        ab = O.relative_to_abs(rel, g['obs_traj'][-1])
        ades.append(O.displacement_error_raw(ab, g['pred_traj_gt']))
        fdes.append(O.final_displacement_error_raw(ab[-1], g['pred_traj_gt'][-1]))
    n = g['obs_traj'].shape[1]
    ade = float(O.best_of_k(ades, g['seq_start_end'])) / (n * cfg['pred_len'])
    fde = float(O.best_of_k(fdes, g['seq_start_end'])) / n
    assert abs(ade - float(g['ade'])) < 1e-5 and abs(fde - float(g['fde'])) < 1e-5


def test_discriminator_matches_reference():
    g = load_golden('discriminator_zara1')
    s = O.discriminator_forward(g['traj'], g['traj_rel'], g['seq_start_end'], state_dict_of(g))
    _close(s, g['scores'], 1e-5)
