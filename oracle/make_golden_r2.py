"""Round-2 golden vectors frozen from the UNMODIFIED reference (run in the build container only):

    python oracle/make_golden_r2.py        # writes tests/golden/{train_step_gat,eval_*_full,pool_g_N,gat_encoder_N,gcn_module_N}.npz

  train_step_gat        one discriminator_step + one generator_step of scripts/train.py:395-484 (imported as they are)
                        on one minibatch, from fixed initial weights, fixed noise (torch.manual_seed before each step)
                        and fixed label-smoothing draws (random.seed): the losses they return, every parameter gradient
                        left in .grad and every updated parameter.
  eval_p_eth_full       the whole ETH test split (obs 8 / pred 8: 195 scenes, 614 peds) through the SGAN-P checkpoint
  eval_gat_zara1_full   the whole zara1 test split (pred 12: 602 scenes, 2253 peds) through the SGAN-GAT checkpoint
                        -- scripts/evaluate_model.py:72-99 with K = 20 and the noise of every (sample, scene) frozen:
                        per-pedestrian raw ADE / FDE of every sample, the best-of-K sums and the ADE / FDE it prints.
TEST INFRASTRUCTURE: nothing in the product reads these files.
"""
import importlib.util
import os
import random
import sys
from types import SimpleNamespace

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle import ref_shim  # noqa: E402
from oracle.make_golden import build_generator, save  # noqa: E402


def _import_script(name):
    cwd = os.getcwd()
    os.chdir(ref_shim.REF_ROOT)            # the scripts do sys.path.append(".")
    try:
        spec = importlib.util.spec_from_file_location('ref_' + name, os.path.join(ref_shim.REF_ROOT, 'scripts', name + '.py'))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        os.chdir(cwd)
    return mod


def make_train_step(R, init_seed=401, write=True):
    train = _import_script('train')
    from sgan.losses import gan_d_loss, gan_g_loss
    ck = ref_shim.load_checkpoint('models/sgan-gat-models/zara1_12_model.pt')
    a = dict(ck['args'])
    args = SimpleNamespace(obs_len=8, pred_len=12, best_k=3, l2_loss_weight=1.0, clipping_threshold_g=2.0,
                           clipping_threshold_d=0.0, g_learning_rate=1e-3, d_learning_rate=1e-3)
    torch.manual_seed(init_seed)
    gen = build_generator(R, a, 'gat')
    disc = R.TrajectoryDiscriminator(obs_len=8, pred_len=12, embedding_dim=a['embedding_dim'], h_dim=a['encoder_h_dim_d'],
                                     mlp_dim=a['mlp_dim'], num_layers=a['num_layers'], dropout=a['dropout'],
                                     batch_norm=a['batch_norm'], d_type=a['d_type'])
    gen.apply(train.init_weights)          # scripts/train.py:194-195, 211-212
    disc.apply(train.init_weights)
    gen.train(); disc.train()
    g0 = {k: v.clone() for k, v in gen.state_dict().items()}
    d0 = {k: v.clone() for k, v in disc.state_dict().items()}
    opt_g = torch.optim.Adam(gen.parameters(), lr=args.g_learning_rate)
    opt_d = torch.optim.Adam(disc.parameters(), lr=args.d_learning_rate)
    _, loader = ref_shim.load_dataset('zara1', 'test', obs_len=8, pred_len=12, batch_size=8)
    batch = next(iter(loader))
    random.seed(77)                        # label smoothing: sgan/losses.py:32,45-46 draw from the global RNG
    torch.manual_seed(500)
    losses_d = train.discriminator_step(args, batch, gen, disc, gan_d_loss, opt_d)
    d_grads = {k: p.grad.clone() for k, p in disc.named_parameters() if p.grad is not None}
    torch.manual_seed(501)
    losses_g = train.generator_step(args, batch, gen, disc, gan_g_loss, opt_g)
    g_grads = {k: p.grad.clone() for k, p in gen.named_parameters() if p.grad is not None}
    if not write:
        return losses_d, losses_g
    names = ('obs_traj', 'pred_traj_gt', 'obs_traj_rel', 'pred_traj_gt_rel', 'obs_traj_rel_v', 'pred_traj_rel_v',
             'obs_traj_g', 'pred_traj_g', 'non_linear_ped', 'loss_mask', 'seq_start_end')
    arrays = {'batch.' + n: t for n, t in zip(names, batch)}
    arrays.update({'g0.' + k: v for k, v in g0.items()})
    arrays.update({'d0.' + k: v for k, v in d0.items()})
    arrays.update({'g1.' + k: v for k, v in gen.state_dict().items()})
    arrays.update({'d1.' + k: v for k, v in disc.state_dict().items()})
    arrays.update({'ggrad.' + k: v for k, v in g_grads.items()})
    arrays.update({'dgrad.' + k: v for k, v in d_grads.items()})
    arrays.update({'loss.' + k: float(v) for k, v in {**losses_d, **losses_g}.items()})
    save('train_step_gat', best_k=args.best_k, l2_loss_weight=args.l2_loss_weight, clipping_threshold_g=args.clipping_threshold_g,
         clipping_threshold_d=args.clipping_threshold_d, g_learning_rate=args.g_learning_rate,
         d_learning_rate=args.d_learning_rate, init_seed=init_seed, label_seed=77, seed_d=500, seed_g=501, alpha=a.get('alpha', 0.2), **arrays)
    print('train step losses', losses_d, losses_g)


def make_full_split(R, name, ckpt_rel, wiring, dset, k_samples, seed, batch_scenes=64):
    ck = ref_shim.load_checkpoint(ckpt_rel)
    args = dict(ck['args'])
    torch.manual_seed(seed)
    g = build_generator(R, args, wiring)
    missing, unexpected = g.load_state_dict(ck['g_state'], strict=False)
    assert all(k.startswith(('gatencoder', 'gcn_module')) for k in missing), missing
    g.train()                              # scripts/evaluate_model.py:54
    _, loader = ref_shim.load_dataset(dset, 'test', obs_len=args['obs_len'], pred_len=args['pred_len'], batch_size=batch_scenes)
    from sgan.losses import displacement_error, final_displacement_error
    from sgan.utils import relative_to_abs
    parts = {k: [] for k in ('obs_traj', 'obs_traj_rel', 'obs_traj_g', 'pred_traj_gt')}
    sse_all, noise_all, ade_raw, fde_raw, rel01 = [], [], [], [], []
    ade_outer, fde_outer, off, total = [], [], 0, 0
    gen_noise = torch.Generator().manual_seed(seed + 1)
    with torch.no_grad():
        for batch in loader:               # the loop of scripts/evaluate_model.py:75-96
            (obs_traj, pred_traj_gt, obs_traj_rel, _pr, _ov, _pv, obs_traj_g, _pg, _nl, _lm, sse) = batch
            noise = torch.randn(k_samples, sse.shape[0], args['noise_dim'][0], generator=gen_noise)
            ade, fde = [], []
            total += pred_traj_gt.size(1)
            for k in range(k_samples):
                rel = g(obs_traj, obs_traj_rel, sse, obs_traj_g, user_noise=noise[k])
                ab = relative_to_abs(rel, obs_traj[-1])
                ade.append(displacement_error(ab, pred_traj_gt, mode='raw'))
                fde.append(final_displacement_error(ab[-1], pred_traj_gt[-1], mode='raw'))
                if k < 2:
                    rel01.append((k, rel))
            # evaluate_helper, scripts/evaluate_model.py:58-69
            a_st, f_st = torch.stack(ade, 1), torch.stack(fde, 1)
            ade_outer.append(sum(torch.min(a_st[s:e].sum(0)) for s, e in sse.tolist()))
            fde_outer.append(sum(torch.min(f_st[s:e].sum(0)) for s, e in sse.tolist()))
            ade_raw.append(a_st); fde_raw.append(f_st)
            for key, t in zip(('obs_traj', 'obs_traj_rel', 'obs_traj_g', 'pred_traj_gt'),
                              (obs_traj, obs_traj_rel, obs_traj_g, pred_traj_gt)):
                parts[key].append(t)
            sse_all.append(sse + off)
            off += obs_traj.shape[1]
            noise_all.append(noise)
    ade = float(sum(ade_outer)) / (total * args['pred_len'])
    fde = float(sum(fde_outer)) / total
    sd = {k: v for k, v in g.state_dict().items() if not k.startswith('gatencoder.fn')}
    if wiring != 'gat':
        sd = {k: v for k, v in sd.items() if not k.startswith('gatencoder.')}
    rel_k0 = torch.cat([r for k, r in rel01 if k == 0], dim=1)
    save(name, seq_start_end=torch.cat(sse_all, 0), noise=torch.cat(noise_all, 1), ade_raw=torch.cat(ade_raw, 0),
         fde_raw=torch.cat(fde_raw, 0), pred_rel_k0=rel_k0, ade=ade, fde=fde, ade_sum=float(sum(ade_outer)),
         fde_sum=float(sum(fde_outer)), pred_len=args['pred_len'], wiring=wiring, batch_scenes=batch_scenes,
         alpha=args.get('alpha', 0.2), n_heads=args.get('n_heads', 1),
         **{k: torch.cat(v, 1) for k, v in parts.items()}, **{'sd.' + k: v for k, v in sd.items()})
    print(name, 'scenes', int(torch.cat(sse_all, 0).shape[0]), 'peds', total, 'ADE %.4f FDE %.4f' % (ade, fde))


def main():
    R = ref_shim.load()
    # A freshly initialised discriminator often scores every fake trajectory 0 (its classifier ends in a ReLU,
    # sgan/models.py:964-970), which kills the adversarial gradient into G: pick the first init seed whose fake scores
    # are alive so that the golden exercises the whole G <- D path.
    seed = 401
    while True:
        ld, lg = make_train_step(R, seed, write=False)
        if abs(lg['G_discriminator_loss'] - 0.6931472) > 0.02 and abs(ld['D_data_loss'] - 1.3862944) > 0.02:
            break
        seed += 1
    make_train_step(R, seed)
    make_full_split(R, 'eval_p_eth_full', 'models/sgan-p-models/eth_8_model.pt', 'mlp', 'eth', 20, seed=601)
    make_full_split(R, 'eval_gat_zara1_full', 'models/sgan-gat-models/zara1_12_model.pt', 'gat', 'zara1', 20, seed=602)
    make_dense(R)


if __name__ == '__main__':
    main()


def make_dense(R):
    """Dense-crowd fixtures (BASELINE.json configs[3]: scenes of 64-1024 pedestrians): the reference modules on ONE large
    scene -- PoolHiddenNet materialises N^2 x 512, GATEncoder [N,N,144] (604 MB at N = 1024).  Forward outputs only."""
    from oracle.make_golden import labels_for, sse_from_sizes, sd_arrays
    for n in (256, 1024):
        torch.manual_seed(700 + n)
        rng = np.random.RandomState(700 + n)
        sse = sse_from_sizes([n])
        with torch.no_grad():
            m = R.PoolHiddenNet(embedding_dim=16, h_dim=32, mlp_dim=64, bottleneck_dim=8, batch_norm=False)
            h = torch.randn(1, n, 32)
            pos = torch.rand(n, 2) * 15
            save('pool_g_%d' % n, seq_start_end=sse, h=h, pos=pos, out=m(h, sse, pos), **sd_arrays(m))
            x = torch.randn(n, 40)
            lab = labels_for([n], rng, p_zero=0.1)
            gat = R.GATEncoder(n_units=None, n_heads=1, dropout=0, alpha=0.2)
            save('gat_encoder_%d' % n, seq_start_end=sse, x=x, pos=pos, labels=lab, out=gat(x, sse, pos, lab), n_heads=1,
                 alpha=0.2, **sd_arrays(gat))
            gcn = R.GCNModule(input_dim=40, hidden_dim=72, out_dim=16, gcn_layers=2, final_dim=24)
            for p in gcn.parameters():
                if p.dim() == 2 and p.shape[0] in (40, 72, 16) and p.shape[1] in (72, 16):
                    p.mul_(0.15)
            save('gcn_module_%d' % n, seq_start_end=sse, x=x, pos=pos, labels=lab, out=gcn(x, sse, pos, lab),
                 **sd_arrays(gcn))
