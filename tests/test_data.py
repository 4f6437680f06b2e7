"""Input pipeline: group_gan_gcn_gat_b200.data against golden vectors made by the unmodified
reference dataset class (oracle/make_golden_data.py), plus host-side properties that need no reference."""
import os
import types

import numpy as np
import pytest
import torch
from torch.utils.data import DataLoader

from group_gan_gcn_gat_b200 import data as D

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
CASES = {'a': dict(obs_len=8, pred_len=12, skip=1, seed=11), 'b': dict(obs_len=4, pred_len=4, skip=2, seed=12)}
ATTRS = ('obs_traj', 'pred_traj', 'obs_traj_rel', 'pred_traj_rel', 'obs_traj_g', 'pred_traj_g', 'loss_mask',
         'non_linear_ped')


@pytest.fixture(scope='module')
def golden():
    with np.load(os.path.join(GOLD, 'dataset_small.npz')) as z:
        return {k: z[k] for k in z.files}


def _dataset(name, **kw):
    c = CASES[name]
    return D.TrajectoryDataset(os.path.join(GOLD, 'data_small_%s' % name), obs_len=c['obs_len'], pred_len=c['pred_len'],
                               skip=c['skip'], **kw)


@pytest.mark.parametrize('name', sorted(CASES))
def test_dataset_tensors_bit_equal_reference(golden, name):
    ds = _dataset(name)
    for a in ATTRS:
        got, ref = getattr(ds, a), golden['%s.%s' % (name, a)]
        assert got.dtype == torch.float32 and tuple(got.shape) == ref.shape, a
        assert np.array_equal(got.numpy(), ref), a
    assert np.array_equal(np.asarray(ds.seq_start_end, dtype=np.int64), golden['%s.seq_start_end' % name])
    assert len(ds) == ds.num_seq == len(ds.seq_start_end)
    assert all(e - s > 1 for s, e in ds.seq_start_end)            # min_ped = 1: at least two pedestrians per sequence


@pytest.mark.parametrize('name', sorted(CASES))
@pytest.mark.parametrize('loader', ['DataLoader', 'DeviceLoader'])
def test_batches_match_reference_loader_under_same_seed(golden, name, loader):
    ds = _dataset(name)
    torch.manual_seed(100 + CASES[name]['seed'])
    it = DataLoader(ds, batch_size=8, shuffle=True, collate_fn=D.seq_collate) if loader == 'DataLoader' else \
        D.DeviceLoader(ds, batch_size=8, shuffle=True)
    for b, batch in enumerate(it):
        if b >= 2:
            break
        assert len(batch) == len(D.BATCH_FIELDS) == 11
        for k, t in enumerate(batch):
            ref = golden['%s.batch%d.%d' % (name, b, k)]
            assert tuple(t.shape) == ref.shape and np.array_equal(t.numpy(), ref), (b, D.BATCH_FIELDS[k])
        assert batch[-1].dtype == torch.int64


def test_device_loader_host_mode_equals_collate_of_items():
    ds = _dataset('a')
    for shuffle, drop_last in ((False, False), (True, True)):
        torch.manual_seed(3)
        ref = list(DataLoader(ds, batch_size=6, shuffle=shuffle, drop_last=drop_last, collate_fn=D.seq_collate))
        torch.manual_seed(3)
        loader = D.DeviceLoader(ds, batch_size=6, shuffle=shuffle, drop_last=drop_last)
        got = list(loader)
        assert len(got) == len(ref) == len(loader)
        for x, y in zip(got, ref):
            assert all(torch.equal(p, q) for p, q in zip(x, y))
    batch = ds.collate_indices([4])                                   # single-sequence batch
    assert batch[-1].tolist() == [[0, ds.seq_start_end[4][1] - ds.seq_start_end[4][0]]]
    assert torch.equal(batch[4], batch[2] * 2.5)                      # velocity = displacement / 0.4 s


def test_cache_round_trip(tmp_path):
    first = _dataset('b', cache_dir=str(tmp_path))
    files = os.listdir(tmp_path)
    assert len(files) == 1 and files[0].startswith('trajectories_')
    again = _dataset('b', cache_dir=str(tmp_path))                    # served from the cache
    for a in ATTRS:
        assert torch.equal(getattr(first, a), getattr(again, a))
    assert first.seq_start_end == again.seq_start_end
    other = D.TrajectoryDataset(os.path.join(GOLD, 'data_small_b'), obs_len=4, pred_len=4, skip=1, cache_dir=str(tmp_path))
    assert len(os.listdir(tmp_path)) == 2 and len(other) != len(first)   # different arguments, different key


def test_unlabelled_files_raise_like_the_reference(tmp_path):
    rows = np.loadtxt(os.path.join(GOLD, 'data_small_b', 'scene_b.txt'), delimiter='\t')
    np.savetxt(tmp_path / 'nolabel.txt', rows[:, :4], delimiter='\t', fmt='%.4f')
    with pytest.raises(AssertionError, match='dataset has no labeling'):
        D.TrajectoryDataset(str(tmp_path), obs_len=4, pred_len=4)


def test_poly_fit_and_empty_window_edge_cases(tmp_path):
    t = np.arange(12.0)
    assert D.poly_fit(np.stack([t, 0.5 * t * t]), 12, 0.002) == 0.0                  # exactly quadratic
    assert D.poly_fit(np.stack([t, np.sin(t)]), 12, 0.002) == 1.0
    short = np.array([[0.0, 1, 0, 0, 1], [10.0, 1, 1, 1, 1], [10.0, 2, 2, 2, 1]])
    assert len(D.scan_file(short, 8, 12)['counts']) == 0                           # fewer frames than one window
    np.savetxt(tmp_path / 'short.txt', short, delimiter='\t', fmt='%.1f')
    with pytest.raises(ValueError):                                                   # the reference dies in np.concatenate
        D.TrajectoryDataset(str(tmp_path))


def test_data_loader_signature_matches_reference():
    args = types.SimpleNamespace(obs_len=8, pred_len=12, skip=1, delim='\t', batch_size=64, loader_num_workers=0)
    dset, loader = D.data_loader(args, os.path.join(GOLD, 'data_small_a'))
    batch = next(iter(loader))
    assert len(batch) == 11 and batch[0].shape[0] == 8 and batch[1].shape[0] == 12
    assert batch[-1][-1, 1].item() == batch[0].shape[1] and len(dset) == 35


@pytest.mark.skipif(not os.path.isdir('/root/reference/datasets_group'), reason='reference tree absent')
@pytest.mark.parametrize('split,pred_len', [('eth/test', 8), ('zara1/test', 12)])
def test_real_splits_bit_equal_live_reference(split, pred_len):
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(GOLD), '..', 'oracle'))
    import ref_shim
    ref_shim.load()
    from sgan.data.trajectories_GCN import TrajectoryDataset as Ref
    path = os.path.join('/root/reference/datasets_group', split)
    ref, got = Ref(path, obs_len=8, pred_len=pred_len), D.TrajectoryDataset(path, obs_len=8, pred_len=pred_len)
    for a in ATTRS:
        assert torch.equal(getattr(ref, a), getattr(got, a)), a
    assert ref.seq_start_end == got.seq_start_end


def _eval_noise(batch_index, n_scenes, k_samples=4, dim=8):          # oracle/make_golden_data.py:eval_noise
    return torch.randn(k_samples, n_scenes, dim, generator=torch.Generator().manual_seed(1000 + batch_index))


def _zara1_weights():
    with np.load(os.path.join(GOLD, 'generator_gat_zara1.npz')) as z:
        return {k[3:]: torch.from_numpy(z[k]) for k in z.files if k.startswith('sd.')}


def test_dataset_to_ade_fde_on_the_oracle_matches_reference_evaluation(golden):
    """files -> TrajectoryDataset -> batches -> oracle generator -> best-of-K ADE/FDE == the reference's evaluate()"""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(GOLD), '..'))
    from oracle import sgan_oracle as O
    sd = _zara1_weights()
    cfg = dict(pred_len=12, wiring='gat', pooling=True, pool_every_timestep=False, alpha=0.2, n_heads=1)
    ds = _dataset('a')
    ade_sum = fde_sum = 0.0
    total = 0
    for b, batch in enumerate(D.DeviceLoader(ds, batch_size=16, shuffle=False)):
        obs, gt, obs_rel, grp, sse = batch[0], batch[1], batch[2], batch[6], batch[10]
        noise = _eval_noise(b, sse.shape[0])
        ades, fdes = [], []
        for k in range(4):
            ab = O.relative_to_abs(O.generator_forward(obs, obs_rel, sse, grp, sd, cfg, noise[k]), obs[-1])
            ades.append(O.displacement_error_raw(ab, gt))
            fdes.append(O.final_displacement_error_raw(ab[-1], gt[-1]))
        ade_sum += float(O.best_of_k(ades, sse))
        fde_sum += float(O.best_of_k(fdes, sse))
        total += gt.shape[1]
    assert total == int(golden['eval.total_traj'])
    assert abs(ade_sum / (total * 12) - float(golden['eval.ade'])) < 1e-5
    assert abs(fde_sum / total - float(golden['eval.fde'])) < 1e-5


@pytest.mark.gpu
def test_evaluate_over_device_loader_matches_reference_evaluation(golden):
    """the whole evaluation script on the GPU: prefetching loader -> sgx generator -> fused best-of-K metrics, against
    the numbers the unmodified reference prints for the same files, weights and noise (ADE/FDE tolerance 1e-4)."""
    from group_gan_gcn_gat_b200 import evaluate as E, models as MD
    torch.backends.cudnn.allow_tf32 = False
    gen = MD.TrajectoryGenerator(obs_len=8, pred_len=12, embedding_dim=16, encoder_h_dim=32, decoder_h_dim=32, mlp_dim=64,
                                 noise_dim=(8,), noise_mix_type='global', pooling_type='pool_net',
                                 pool_every_timestep=False, bottleneck_dim=8, batch_norm=False, n_heads=1, alpha=0.2)
    gen.load_state_dict(_zara1_weights(), strict=True)
    gen = gen.cuda().train()
    ds = _dataset('a')
    for hoist, fold in ((False, False), (True, False), (False, True)):     # sample loop, hoisted context, folded samples
        loader = D.DeviceLoader(ds, batch_size=16, shuffle=False, device='cuda:0')
        ade, fde = E.evaluate(dict(pred_len=12), loader, gen, 4, noise_for_batch=_eval_noise, hoist_context=hoist,
                              fold_samples=fold)
        assert abs(float(ade) - float(golden['eval.ade'])) < 1e-4, (hoist, fold, float(ade))
        assert abs(float(fde) - float(golden['eval.fde'])) < 1e-4, (hoist, fold, float(fde))


def test_get_generator_builds_from_checkpoint_args():
    from group_gan_gcn_gat_b200 import evaluate as E
    pytest.importorskip('torch')
    args = dict(obs_len=8, pred_len=12, embedding_dim=16, encoder_h_dim_g=32, decoder_h_dim_g=32, mlp_dim=64, num_layers=1,
                noise_dim=(8,), noise_type='gaussian', noise_mix_type='global', pooling_type='pool_net',
                pool_every_timestep=0, dropout=0, bottleneck_dim=8, neighborhood_size=2.0, grid_size=8, batch_norm=0,
                hidden_units='16', n_heads=1, dropout1=0, alpha=0.2)
    gen = E.get_generator({'args': args, 'g_state': _zara1_weights()}, device='cpu')
    assert gen.training and gen.pred_len == 12
    assert set(gen.state_dict()) == set(_zara1_weights())


@pytest.mark.gpu
def test_device_loader_prefetches_identical_batches_to_the_gpu():
    ds = _dataset('a')
    torch.manual_seed(9)
    host = list(D.DeviceLoader(ds, batch_size=8, shuffle=True))
    torch.manual_seed(9)
    dev = list(D.DeviceLoader(ds, batch_size=8, shuffle=True, device='cuda:0'))
    assert len(host) == len(dev)
    for x, y in zip(host, dev):
        assert all(q.is_cuda and torch.equal(p, q.cpu()) for p, q in zip(x, y))
