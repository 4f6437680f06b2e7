"""Host batches handed to evaluate_batch / evaluate (scripts/evaluate_model.py:72-99 receives its minibatches from a CPU
DataLoader): the staged copies (utils.stage_host_batch: copy stream, one event per tensor, the compute stream waits where
the forward first reads a tensor) must give the results of the same call on device tensors."""
import numpy as np
import pytest
import torch

from conftest import load_golden, state_dict_of
from test_gpu_pdl import _batch, _sgan_p

pytestmark = pytest.mark.gpu

DEV = 'cuda:0'


def _close(a, b, tol=2e-6):
    a, b = float(a), float(b)
    assert abs(a - b) <= tol * max(1.0, abs(b)), (a, b)


@pytest.mark.parametrize('pinned', [True, False])
@pytest.mark.parametrize('mode', ['loop', 'folded', 'hoisted'])
def test_evaluate_batch_from_host_tensors_matches_device_tensors(pinned, mode):
    from group_gan_gcn_gat_b200.evaluate import evaluate_batch
    gen = _sgan_p(DEV)
    n_scenes = 3000
    obs, rel, sse, grp, _ = _batch(n_scenes, 9)
    rng = np.random.RandomState(3)
    gt = torch.from_numpy(rng.randn(8, obs.shape[1], 2).astype(np.float32)).cumsum(0) * 0.3 + obs[-1]
    noise = torch.from_numpy(rng.randn(5, n_scenes, 8).astype(np.float32))
    kw = dict(fold_samples=(mode == 'folded'), hoist_context=(mode == 'hoisted'))
    a0, f0 = evaluate_batch(gen, obs.to(DEV), rel.to(DEV), sse.clone(), grp.to(DEV), gt.to(DEV), 5, noise=noise.to(DEV), **kw)
    host = [t.pin_memory() if pinned else t for t in (obs, rel, grp, gt)]
    for _ in range(3):                       # repeated: the staged blocks are recycled by the caching allocator
        a1, f1 = evaluate_batch(gen, host[0], host[1], sse.clone(), host[2], host[3], 5, noise=noise, **kw)
        _close(a1, a0)
        _close(f1, f0)
    assert a1.is_cuda and torch.isfinite(a1)


def test_evaluate_batch_from_host_tensors_gat_wiring_matches_golden_errors():
    """the SGAN-GAT wiring reads obs_traj_g as well: real checkpoint and data, host tensors in, frozen reference ADE/FDE"""
    import group_gan_gcn_gat_b200.models as MD
    from group_gan_gcn_gat_b200.evaluate import evaluate_batch
    g = load_golden('generator_gat_zara1')
    gen = MD.TrajectoryGenerator(obs_len=8, pred_len=int(g['pred_len']), embedding_dim=16, encoder_h_dim=32,
                                 decoder_h_dim=32, mlp_dim=64, num_layers=1, noise_dim=(8,), noise_type='gaussian',
                                 noise_mix_type='global', pooling_type='pool_net', pool_every_timestep=False, dropout=0,
                                 bottleneck_dim=8, batch_norm=False, n_heads=int(g['n_heads']), dropout1=0,
                                 alpha=float(g['alpha']), context_type='gat')
    gen.load_state_dict(state_dict_of(g), strict=False)
    gen = gen.to(DEV).train()
    k = g['noise'].shape[0]
    for fold in (False, True):
        a, f = evaluate_batch(gen, g['obs_traj'].pin_memory(), g['obs_traj_rel'].pin_memory(), g['seq_start_end'].clone(),
                              g['obs_traj_g'].pin_memory(), g['pred_traj_gt'].pin_memory(), k, noise=g['noise'],
                              fold_samples=fold)
        n = g['obs_traj'].shape[1]
        assert abs(float(a) / (n * int(g['pred_len'])) - float(g['ade'])) < 1e-4
        assert abs(float(f) / n - float(g['fde'])) < 1e-4
