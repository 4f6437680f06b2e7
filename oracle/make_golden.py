"""Freeze golden vectors from the UNMODIFIED reference (run in the build container only).

    python oracle/make_golden.py            # writes tests/golden/*.npz

Every fixture holds seeded inputs, the reference module's parameters (under the
reference's own state_dict names), its outputs and -- for the hot-path modules -- the
gradients of  loss = sum(out * G)  w.r.t. every input and parameter.  The GPU box has no
/root/reference, so these files are what the `-m gpu` parity tests and smoke() check
against.  TEST INFRASTRUCTURE: nothing in the product reads them.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle import ref_shim  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), 'tests', 'golden')


def sse_from_sizes(sizes):
    cs = np.concatenate([[0], np.cumsum(sizes)])
    return torch.tensor([[cs[i], cs[i + 1]] for i in range(len(sizes))], dtype=torch.int64)


def npify(d):
    out = {}
    for k, v in d.items():
        if torch.is_tensor(v):
            v = v.detach().cpu().numpy()
        out[k] = np.asarray(v)
    return out


def save(name, **arrays):
    os.makedirs(OUT, exist_ok=True)
    path = os.path.join(OUT, name + '.npz')
    np.savez_compressed(path, **npify(arrays))
    print('wrote', path, os.path.getsize(path) // 1024, 'KiB')


def sd_arrays(module, tag='sd.'):
    return {tag + k: v for k, v in module.state_dict().items()}


def grads_of(module, out, upstream, inputs):
    loss = (out * upstream).sum()
    params = list(module.named_parameters())
    gs = torch.autograd.grad(loss, [p for _, p in params] + list(inputs.values()), allow_unused=True)
    res = {}
    for (n, _), g in zip(params, gs[:len(params)]):
        res['grad.' + n] = g if g is not None else torch.zeros(())
    for n, g in zip(inputs.keys(), gs[len(params):]):
        res['grad_in.' + n] = g
    return res


def labels_for(sizes, rng, p_zero=0.2):
    labs = []
    for n in sizes:
        hi = max(1, n // 3)
        l = rng.randint(1, hi + 1, size=n).astype(np.float32)
        l[rng.rand(n) < p_zero] = 0.0
        labs.append(l)
    return torch.tensor(np.concatenate(labs)).view(-1, 1)


def make_pool(R, name, e_dim, h_dim, bott, sizes, seed):
    torch.manual_seed(seed)
    m = R.PoolHiddenNet(embedding_dim=e_dim, h_dim=h_dim, mlp_dim=64, bottleneck_dim=bott, batch_norm=False)
    sse = sse_from_sizes(sizes)
    b = int(sse[-1, 1])
    h = torch.randn(1, b, h_dim, requires_grad=True)
    pos = (torch.rand(b, 2) * 15).requires_grad_(True)
    out = m(h, sse, pos)
    up = torch.randn_like(out)
    g = grads_of(m, out, up, {'h': h, 'pos': pos})
    save(name, seq_start_end=sse, h=h, pos=pos, out=out, upstream=up, **sd_arrays(m), **g)


def make_graph_module(R, name, kind, in_dim, n_heads, sizes, seed):
    torch.manual_seed(seed)
    rng = np.random.RandomState(seed)
    if kind == 'gat':
        m = R.GATEncoder(n_units=None, n_heads=n_heads, dropout=0, alpha=0.2)
    else:
        m = R.GCNModule(input_dim=in_dim, hidden_dim=72, out_dim=16, gcn_layers=2, final_dim=24)
        with torch.no_grad():      # plain randn weights blow activations up to 1e3; keep fixtures O(1)
            for p in m.parameters():
                if p.dim() == 2 and p.shape[0] in (in_dim, 72, 16) and p.shape[1] in (72, 16):
                    p.mul_(0.15)
    sse = sse_from_sizes(sizes)
    b = int(sse[-1, 1])
    x = torch.randn(b, in_dim, requires_grad=True)
    pos = torch.rand(b, 2) * 15
    lab = labels_for(sizes, rng)
    out = m(x, sse, pos, lab)
    up = torch.randn_like(out)
    g = grads_of(m, out, up, {'x': x})
    save(name, seq_start_end=sse, x=x, pos=pos, labels=lab, out=out, upstream=up, n_heads=n_heads,
         alpha=0.2, **sd_arrays(m), **g)


def make_dense_layers(R):
    torch.manual_seed(11)
    n = 9
    adj = (torch.rand(n, n) < 0.4).float()
    adj = ((adj + adj.T + torch.eye(n)) > 0).float()
    adj = adj / adj.sum(1, keepdim=True)
    layer = R.GraphAttentionLayer(12, 20, dropout=0, alpha=0.2, concat=True)
    x = torch.randn(n, 12, requires_grad=True)
    out = layer(x, adj)
    up = torch.randn_like(out)
    g = grads_of(layer, out, up, {'x': x})
    save('gat_layer_dense', x=x, adj=adj, out=out, upstream=up, alpha=0.2, **sd_arrays(layer), **g)

    net = R.GAT(12, 20, 6, dropout=0, alpha=0.2, nheads=3)
    x = torch.randn(n, 12, requires_grad=True)
    out = net(x, adj)
    up = torch.randn_like(out)
    g = grads_of(net, out, up, {'x': x})
    save('gat_dense', x=x, adj=adj, out=out, upstream=up, alpha=0.2, n_heads=3, **sd_arrays(net), **g)

    gcn = R.GCN(input_dim=12, hidden_dim=20, out_dim=6, gcn_layers=3)
    with torch.no_grad():
        for p in gcn.parameters():
            p.mul_(0.3)
    x = torch.randn(n, 12, requires_grad=True)
    a = torch.randn(n, n) * 0.3
    out = gcn(a, x)
    up = torch.randn_like(out)
    g = grads_of(gcn, out, up, {'x': x})
    save('gcn_dense', x=x, adj=a, out=out, upstream=up, **sd_arrays(gcn), **g)


def make_groups(R):
    """Dense M_intra / A_intra / R_intra / normalised R for a handful of label vectors."""
    m = R.GCNModule()
    cases = {
        'notebook': [1, 1, 2, 0],                       # Untitled.ipynb:546-556  -> R=[[1,1,0,0],[0,0,1,0],[0,0,0,1]]
        'gcnpy128': [0, 0, 3, 3, 1, 1, 0, 2, 0, 0],     # sgan/GCN.py:128
        'allzero': [0, 0, 0],
        'onegroup': [5, 5, 5, 5, 5],
        'single': [7],
        'frac': [1.5, 2.25, 1.5, 0, 2.25, 2.25, 9],
    }
    arrays = {}
    for k, labs in cases.items():
        g = torch.tensor(labs, dtype=torch.float32).view(-1, 1)
        n = g.shape[0]
        eye = torch.eye(n).bool()
        a_g = g.repeat(1, n)
        b_g = g.transpose(1, 0).repeat(n, 1)
        m_intra = (a_g == b_g) & (a_g != 0) | eye
        a_intra = m.normalize(m_intra, dim=1)
        uniq = torch.unique(m_intra, sorted=False, dim=0)
        rows = [uniq[i].unsqueeze(0) for i in range(uniq.shape[0] - 1, -1, -1)]
        r = torch.cat(rows, dim=0)
        r_n = m.normalize(r, dim=1)
        arrays.update({k + '.labels': g, k + '.M': m_intra, k + '.A': a_intra, k + '.R': r, k + '.Rn': r_n})
    save('groups', **arrays)


class _CtxAdapter(torch.nn.Module):
    """Routes TrajectoryGenerator's context call (sgan/models.py:905) to an older wiring (models.py:898, 902)."""

    def __init__(self, fn):
        super().__init__()
        self.fn = [fn]

    def forward(self, x, sse, pos, grp):
        return self.fn[0](x, sse, pos, grp)


def build_generator(R, args, wiring, pool_every_timestep=None):
    pet = args['pool_every_timestep'] if pool_every_timestep is None else pool_every_timestep
    g = R.TrajectoryGenerator(
        obs_len=args['obs_len'], pred_len=args['pred_len'], embedding_dim=args['embedding_dim'],
        encoder_h_dim=args['encoder_h_dim_g'], decoder_h_dim=args['decoder_h_dim_g'], mlp_dim=args['mlp_dim'],
        num_layers=args['num_layers'], noise_dim=tuple(args['noise_dim']), noise_type=args['noise_type'],
        noise_mix_type=args['noise_mix_type'], pooling_type=args['pooling_type'], pool_every_timestep=pet,
        dropout=args['dropout'], bottleneck_dim=args['bottleneck_dim'], neighborhood_size=args['neighborhood_size'],
        grid_size=args['grid_size'], batch_norm=args['batch_norm'], n_units=None,
        n_heads=args.get('n_heads', 1), dropout1=args.get('dropout1', 0), alpha=args.get('alpha', 0.2))
    if wiring == 'mlp':
        in_dim = args['encoder_h_dim_g'] + (args['bottleneck_dim'] if args['pooling_type'] else 0)
        g.mlp_decoder_context = R.make_mlp([in_dim, args['mlp_dim'], args['decoder_h_dim_g'] - args['noise_dim'][0]],
                                           batch_norm=args['batch_norm'], dropout=args['dropout'])
        ctx = g.mlp_decoder_context
        g.gatencoder = _CtxAdapter(lambda x, sse, pos, grp: ctx(x))
    elif wiring == 'gcn':
        mod = g.gcn_module
        g.gatencoder = _CtxAdapter(lambda x, sse, pos, grp: mod(x, sse, pos, grp))
    return g


def make_generator(R, name, ckpt_rel, wiring, dset, n_scenes, k_samples, seed, pool_every_timestep=None,
                   fresh=False):
    ck = ref_shim.load_checkpoint(ckpt_rel)
    args = dict(ck['args'])
    torch.manual_seed(seed)
    g = build_generator(R, args, wiring, pool_every_timestep)
    if not fresh:
        missing, unexpected = g.load_state_dict(ck['g_state'], strict=False)
        assert all(k.startswith(('gatencoder', 'gcn_module')) for k in missing), missing
        print(name, 'unexpected keys ignored:', unexpected)
    else:
        for mod in g.modules():
            if isinstance(mod, torch.nn.Linear):
                torch.nn.init.kaiming_normal_(mod.weight)
    g.train()
    _, loader = ref_shim.load_dataset(dset, 'test', obs_len=args['obs_len'], pred_len=args['pred_len'], batch_size=n_scenes)
    batch = next(iter(loader))
    (obs_traj, pred_traj_gt, obs_traj_rel, _pr, _ov, _pv, obs_traj_g, _pg, _nl, _lm, sse) = batch
    noise = torch.randn(k_samples, sse.shape[0], args['noise_dim'][0])
    from sgan.utils import relative_to_abs
    from sgan.losses import displacement_error, final_displacement_error
    rels, ades, fdes = [], [], []
    with torch.no_grad():
        for k in range(k_samples):
            rel = g(obs_traj, obs_traj_rel, sse, obs_traj_g, user_noise=noise[k])
            ab = relative_to_abs(rel, obs_traj[-1])
            rels.append(rel)
            ades.append(displacement_error(ab, pred_traj_gt, mode='raw'))
            fdes.append(final_displacement_error(ab[-1], pred_traj_gt[-1], mode='raw'))
    ade_sum = sum(torch.min(torch.stack(ades, 1)[s:e].sum(0)) for s, e in sse.tolist())
    fde_sum = sum(torch.min(torch.stack(fdes, 1)[s:e].sum(0)) for s, e in sse.tolist())
    n_traj = obs_traj.shape[1]
    sd = {k: v for k, v in g.state_dict().items() if not k.startswith('gatencoder.fn')}
    if wiring != 'gat':
        sd = {k: v for k, v in sd.items() if not k.startswith('gatencoder.')}
    save(name, obs_traj=obs_traj, obs_traj_rel=obs_traj_rel, obs_traj_g=obs_traj_g, pred_traj_gt=pred_traj_gt,
         seq_start_end=sse, noise=noise, pred_rel=torch.stack(rels, 0),
         ade=float(ade_sum) / (n_traj * args['pred_len']), fde=float(fde_sum) / n_traj,
         pred_len=args['pred_len'], wiring=wiring,
         pool_every_timestep=int(args['pool_every_timestep'] if pool_every_timestep is None else pool_every_timestep),
         alpha=args.get('alpha', 0.2), n_heads=args.get('n_heads', 1),
         **{'sd.' + k: v for k, v in sd.items()})


def make_discriminator(R, name, ckpt_rel, dset, n_scenes, seed):
    ck = ref_shim.load_checkpoint(ckpt_rel)
    args = dict(ck['args'])
    torch.manual_seed(seed)
    d = R.TrajectoryDiscriminator(obs_len=args['obs_len'], pred_len=args['pred_len'], embedding_dim=args['embedding_dim'],
                                  h_dim=args['encoder_h_dim_d'], mlp_dim=args['mlp_dim'], num_layers=args['num_layers'],
                                  dropout=args['dropout'], batch_norm=args['batch_norm'], d_type=args['d_type'])
    d.load_state_dict(ck['d_state'])
    d.train()
    _, loader = ref_shim.load_dataset(dset, 'test', obs_len=args['obs_len'], pred_len=args['pred_len'], batch_size=n_scenes)
    batch = next(iter(loader))
    (obs_traj, pred_traj_gt, obs_traj_rel, pred_traj_rel, _ov, _pv, _g, _pg, _nl, _lm, sse) = batch
    traj = torch.cat([obs_traj, pred_traj_gt], 0)
    traj_rel = torch.cat([obs_traj_rel, pred_traj_rel], 0)
    with torch.no_grad():
        scores = d(traj, traj_rel, sse)
    save(name, traj=traj, traj_rel=traj_rel, seq_start_end=sse, scores=scores, **sd_arrays(d))


def main():
    R = ref_shim.load()
    make_groups(R)
    make_pool(R, 'pool_g', 16, 32, 8, [2, 3, 1, 7, 13, 4], seed=101)
    make_pool(R, 'pool_d', 16, 48, 48, [2, 5, 9, 3], seed=102)
    make_pool(R, 'pool_g_big', 16, 32, 8, [70, 2, 33], seed=103)
    make_graph_module(R, 'gat_encoder_h1', 'gat', 40, 1, [2, 3, 1, 7, 13, 4, 6], seed=201)
    make_graph_module(R, 'gat_encoder_h2', 'gat', 40, 2, [5, 2, 9], seed=202)
    make_graph_module(R, 'gcn_module_40', 'gcn', 40, 0, [2, 3, 1, 7, 13, 4, 6], seed=203)
    make_graph_module(R, 'gcn_module_32', 'gcn', 32, 0, [5, 2, 9], seed=204)
    make_dense_layers(R)
    make_generator(R, 'generator_gat_zara1', 'models/sgan-gat-models/zara1_12_model.pt', 'gat', 'zara1', 6, 3, seed=301)
    make_generator(R, 'generator_p_eth', 'models/sgan-p-models/eth_8_model.pt', 'mlp', 'eth', 6, 3, seed=302)
    make_generator(R, 'generator_gcn_zara1', 'models/sgan-g-p-models/zara1_12_model.pt', 'gcn', 'zara1', 6, 2, seed=303)
    make_generator(R, 'generator_gat_pet', 'models/sgan-gat-models/zara1_12_model.pt', 'gat', 'zara1', 4, 2, seed=304,
                   pool_every_timestep=1, fresh=True)
    make_discriminator(R, 'discriminator_zara1', 'models/sgan-gat-models/zara1_12_model.pt', 'zara1', 6, seed=305)


if __name__ == '__main__':
    main()
