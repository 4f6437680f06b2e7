"""Training path of the fused recurrences (sgx_lstm_*_train_fwd / sgx_lstm_bwd) against a plain PyTorch fp32
reference of the same computation: nn.LSTM (cuDNN, TF32 off) + the step loop of sgan/models.py:62-92, 142-178."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu


def _close(a, b, tol, what):
    scale = max(float(b.detach().abs().max()), 1e-6)
    err = float((a.detach() - b.detach()).abs().max()) / scale
    assert err < tol, '%s: rel err %.3e (scale %.3e)' % (what, err, scale)


def _mods(E, H, seed):
    torch.manual_seed(seed)
    emb = torch.nn.Linear(2, E).cuda()
    lstm = torch.nn.LSTM(E, H, 1).cuda()
    hp = torch.nn.Linear(H, 2).cuda()
    return emb, lstm, hp


def _grads(outputs, weights, inputs, params):
    loss = sum((o * w).sum() for o, w in zip(outputs, weights))
    return torch.autograd.grad(loss, list(inputs) + list(params), allow_unused=True)


@pytest.mark.parametrize('H,T,batch', [(32, 8, 300), (48, 20, 517), (64, 3, 1)])
def test_encoder_train_matches_autograd_reference(H, T, batch):
    from group_gan_gcn_gat_b200 import ops
    torch.backends.cudnn.allow_tf32 = False
    E = 16
    emb, lstm, _ = _mods(E, H, 3)
    g = torch.Generator(device='cuda').manual_seed(1)
    seq = (torch.randn(T, batch, 2, device='cuda', generator=g) * 0.5).requires_grad_(True)
    w_out = torch.randn(1, batch, H, device='cuda', generator=g)
    params = list(emb.parameters()) + list(lstm.parameters())
    ref_h = lstm(emb(seq.reshape(-1, 2)).view(T, batch, E))[1][0]
    got_h = ops.lstm_encoder_train(seq, emb, lstm)
    _close(got_h, ref_h, 2e-5, 'final_h')
    ref_g = _grads([ref_h], [w_out], [seq], params)
    got_g = _grads([got_h], [w_out], [seq], params)
    names = ['d_seq', 'We', 'be', 'W_ih', 'W_hh', 'b_ih', 'b_hh']
    for n, a, b in zip(names, got_g, ref_g):
        _close(a, b, 2e-4, n)


@pytest.mark.parametrize('H,steps,batch,with_c0', [(32, 12, 257, False), (32, 8, 64, True), (48, 5, 130, True)])
def test_decoder_train_matches_autograd_reference(H, steps, batch, with_c0):
    from group_gan_gcn_gat_b200 import ops
    torch.backends.cudnn.allow_tf32 = False
    E = 16
    emb, lstm, hp = _mods(E, H, 5)
    g = torch.Generator(device='cuda').manual_seed(2)
    h0 = torch.randn(batch, H, device='cuda', generator=g).requires_grad_(True)
    c0 = (torch.randn(batch, H, device='cuda', generator=g) * 0.5).requires_grad_(True) if with_c0 else None
    rel0 = torch.randn(batch, 2, device='cuda', generator=g) * 0.3
    w_pred = torch.randn(steps, batch, 2, device='cuda', generator=g)
    w_h = torch.randn(batch, H, device='cuda', generator=g)
    params = list(emb.parameters()) + list(lstm.parameters()) + list(hp.parameters())

    def reference():
        state = (h0.unsqueeze(0), (c0 if with_c0 else torch.zeros_like(h0)).unsqueeze(0))
        x = emb(rel0).view(1, batch, E)
        out = []
        for _ in range(steps):
            o, state = lstm(x, state)
            rel = hp(o.view(-1, H))
            x = emb(rel).view(1, batch, E)
            out.append(rel)
        return torch.stack(out, 0), state[0][0]

    ref_pred, ref_h = reference()
    got_pred, got_h = ops.lstm_decoder_train(h0, c0, rel0, steps, emb, lstm, hp)
    _close(got_pred, ref_pred, 3e-5, 'pred_rel')
    _close(got_h, ref_h, 3e-5, 'final_h')
    inputs = [h0] + ([c0] if with_c0 else [])
    ref_g = _grads([ref_pred, ref_h], [w_pred, w_h], inputs, params)
    got_g = _grads([got_pred, got_h], [w_pred, w_h], inputs, params)
    names = ['d_h0'] + (['d_c0'] if with_c0 else []) + ['We', 'be', 'W_ih', 'W_hh', 'b_ih', 'b_hh', 'W_hp', 'b_hp']
    for n, a, b in zip(names, got_g, ref_g):
        _close(a, b, 3e-4, n)


def test_generator_and_discriminator_use_the_fused_training_recurrences(monkeypatch):
    """one G forward/backward and one D forward/backward with the fused path on and off give the same gradients"""
    from group_gan_gcn_gat_b200 import models as MD
    torch.backends.cudnn.allow_tf32 = False
    torch.manual_seed(11)
    gen = MD.TrajectoryGenerator(obs_len=8, pred_len=12, embedding_dim=16, encoder_h_dim=32, decoder_h_dim=32, mlp_dim=64,
                                 noise_dim=(8,), noise_mix_type='global', pooling_type='pool_net',
                                 pool_every_timestep=False, bottleneck_dim=8, batch_norm=False, n_heads=1).cuda()
    disc = MD.TrajectoryDiscriminator(obs_len=8, pred_len=12, embedding_dim=16, h_dim=48, mlp_dim=64, batch_norm=False,
                                      d_type='global').cuda()
    sizes = [3, 5, 2, 9, 4]
    n = sum(sizes)
    ends = torch.tensor(sizes).cumsum(0)
    sse = torch.stack([ends - torch.tensor(sizes), ends], 1).cuda()
    obs_rel = torch.randn(8, n, 2, device='cuda') * 0.3
    obs = torch.randn(1, n, 2, device='cuda') * 5 + obs_rel.cumsum(0)
    grp = torch.randint(0, 3, (8, n, 1), device='cuda').float()
    noise = torch.randn(len(sizes), 8, device='cuda')

    def run(flag):
        monkeypatch.setenv('SGX_LSTM_TRAIN', flag)
        gen.zero_grad(); disc.zero_grad()
        rel = gen(obs, obs_rel, sse, grp, user_noise=noise)
        traj_rel = torch.cat([obs_rel, rel], 0)
        traj = torch.cat([obs, obs[-1:] + rel.cumsum(0)], 0)
        scores = disc(traj, traj_rel, sse)
        (scores.sum() + (rel ** 2).sum()).backward()
        return rel.detach(), scores.detach(), [p.grad.clone() for p in list(gen.parameters()) + list(disc.parameters())
                                               if p.grad is not None]

    rel1, sc1, g1 = run('1')
    rel0, sc0, g0 = run('0')
    _close(rel1, rel0, 5e-5, 'pred_rel')
    _close(sc1, sc0, 5e-5, 'scores')
    assert len(g1) == len(g0) > 20
    biggest = max(float(b.abs().max()) for b in g0)
    for k, (a, b) in enumerate(zip(g1, g0)):
        err = float((a - b).abs().max()) / max(float(b.abs().max()), 1e-3 * biggest)
        assert err < 2e-3, 'grad %d: %.3e' % (k, err)


def test_generator_step_with_folded_samples_matches_the_sample_loop():
    """parallel.generator_step: best_k samples as one forward over best_k copies of the batch == the reference's loop
    (scripts/train.py:443-455) -- same noise stream, same losses, same parameter update."""
    import copy
    from types import SimpleNamespace
    from group_gan_gcn_gat_b200 import models as MD, parallel
    torch.backends.cudnn.allow_tf32 = False
    torch.manual_seed(21)
    gen = MD.TrajectoryGenerator(obs_len=8, pred_len=12, embedding_dim=16, encoder_h_dim=32, decoder_h_dim=32, mlp_dim=64,
                                 noise_dim=(8,), noise_mix_type='global', pooling_type='pool_net',
                                 pool_every_timestep=False, bottleneck_dim=8, batch_norm=False, n_heads=1).cuda()
    disc = MD.TrajectoryDiscriminator(obs_len=8, pred_len=12, embedding_dim=16, h_dim=48, mlp_dim=64, batch_norm=False,
                                      d_type='global').cuda()
    with torch.no_grad():
        for p in gen.gcn_module.parameters():
            p.mul_(0.1)
    sizes = [4, 2, 7, 3, 5, 2]
    n = sum(sizes)
    ends = torch.tensor(sizes).cumsum(0)
    sse = torch.stack([ends - torch.tensor(sizes), ends], 1).cuda()
    rel = torch.randn(20, n, 2, device='cuda') * 0.3
    rel[0] = 0
    traj = torch.randn(1, n, 2, device='cuda') * 5 + rel.cumsum(0)
    grp = torch.randint(0, 3, (1, n, 1), device='cuda').float().expand(8, n, 1).contiguous()
    batch = (traj[:8], traj[8:], rel[:8], rel[8:], grp, torch.ones(n, 20, device='cuda'), sse)
    out = {}
    for fold in (True, False):
        g, d = copy.deepcopy(gen), copy.deepcopy(disc)
        opt = torch.optim.Adam(g.parameters(), lr=1e-3)
        args = SimpleNamespace(obs_len=8, pred_len=12, best_k=4, l2_loss_weight=1.0, clipping_threshold_g=2.0,
                               fold_best_k=fold)
        torch.manual_seed(77)
        losses = parallel.generator_step(args, batch, g, d, opt, label_rng=parallel.make_label_rng(0, 0))
        out[fold] = (losses, [p.detach().clone() for p in g.parameters()])
    for k in out[True][0]:
        assert abs(float(out[True][0][k]) - float(out[False][0][k])) < 1e-4 * max(1.0, abs(float(out[False][0][k]))), k
    moved = 0
    for a, b, p0 in zip(out[True][1], out[False][1], gen.parameters()):
        moved += int(not torch.equal(b, p0.detach()))
        assert float((a - b).abs().max()) < 2e-5          # Adam step = lr * sign-like update, lr = 1e-3
    assert moved > 20


def test_discriminator_step_stacked_batch_matches_two_calls():
    """parallel.discriminator_step (generator under no_grad, fake + real as one D batch) == the reference's two D calls
    with the generator graph attached (scripts/train.py:395-429): same loss, same D update."""
    import copy
    from types import SimpleNamespace
    from group_gan_gcn_gat_b200 import models as MD, parallel
    from group_gan_gcn_gat_b200.losses import gan_d_loss
    from group_gan_gcn_gat_b200.utils import relative_to_abs
    torch.backends.cudnn.allow_tf32 = False
    torch.manual_seed(31)
    gen = MD.TrajectoryGenerator(obs_len=8, pred_len=12, embedding_dim=16, encoder_h_dim=32, decoder_h_dim=32, mlp_dim=64,
                                 noise_dim=(8,), noise_mix_type='global', pooling_type='pool_net',
                                 pool_every_timestep=False, bottleneck_dim=8, batch_norm=False, n_heads=1).cuda()
    disc = MD.TrajectoryDiscriminator(obs_len=8, pred_len=12, embedding_dim=16, h_dim=48, mlp_dim=64, batch_norm=False,
                                      d_type='global').cuda()
    sizes = [3, 6, 2, 4]
    n = sum(sizes)
    ends = torch.tensor(sizes).cumsum(0)
    sse = torch.stack([ends - torch.tensor(sizes), ends], 1).cuda()
    rel = torch.randn(20, n, 2, device='cuda') * 0.3
    rel[0] = 0
    traj = torch.randn(1, n, 2, device='cuda') * 5 + rel.cumsum(0)
    grp = torch.randint(0, 3, (1, n, 1), device='cuda').float().expand(8, n, 1).contiguous()
    batch = (traj[:8], traj[8:], rel[:8], rel[8:], grp, torch.ones(n, 20, device='cuda'), sse)
    args = SimpleNamespace(obs_len=8, pred_len=12, clipping_threshold_d=0.0)
    # ours
    d1 = copy.deepcopy(disc)
    torch.manual_seed(5)
    out = parallel.discriminator_step(args, batch, gen, d1, torch.optim.Adam(d1.parameters(), lr=1e-3),
                                      label_rng=parallel.make_label_rng(0, 0))
    # the reference's sequence, spelled out
    d2 = copy.deepcopy(disc)
    opt = torch.optim.Adam(d2.parameters(), lr=1e-3)
    torch.manual_seed(5)
    fake_rel = gen(batch[0], batch[2], sse, grp)
    fake = relative_to_abs(fake_rel, batch[0][-1])
    s_fake = d2(torch.cat([batch[0], fake], 0), torch.cat([batch[2], fake_rel], 0), sse)
    s_real = d2(torch.cat([batch[0], batch[1]], 0), torch.cat([batch[2], batch[3]], 0), sse)
    loss = gan_d_loss(s_real, s_fake, parallel.make_label_rng(0, 0))
    opt.zero_grad()
    loss.backward()
    opt.step()
    assert abs(float(out['D_total_loss']) - float(loss)) < 1e-5 * max(1.0, abs(float(loss)))
    for a, b in zip(d1.parameters(), d2.parameters()):
        assert float((a - b).abs().max()) < 2e-5


def test_cuda_graph_replay_of_the_generator_matches_eager():
    """evaluate.GraphedGenerator: every launch of a forward is capture-safe; replays with new inputs / noise are
    bit-identical to eager calls, a new scene layout re-captures."""
    import numpy as np
    from group_gan_gcn_gat_b200 import models as MD
    from group_gan_gcn_gat_b200.evaluate import GraphedGenerator
    torch.backends.cudnn.allow_tf32 = False
    torch.manual_seed(41)
    gen = MD.TrajectoryGenerator(obs_len=8, pred_len=12, embedding_dim=16, encoder_h_dim=32, decoder_h_dim=32, mlp_dim=64,
                                 noise_dim=(8,), noise_mix_type='global', pooling_type='pool_net',
                                 pool_every_timestep=False, bottleneck_dim=8, batch_norm=False, n_heads=1).cuda()
    graphed = GraphedGenerator(gen)

    def batch(sizes, seed):
        g = torch.Generator(device='cuda').manual_seed(seed)
        n = sum(sizes)
        ends = torch.tensor(sizes).cumsum(0)
        sse = torch.stack([ends - torch.tensor(sizes), ends], 1).cuda()
        rel = torch.randn(8, n, 2, device='cuda', generator=g) * 0.3
        obs = torch.randn(1, n, 2, device='cuda', generator=g) * 5 + rel.cumsum(0)
        grp = torch.randint(0, 3, (8, n, 1), device='cuda', generator=g).float()
        return obs, rel, sse, grp, torch.randn(len(sizes), 8, device='cuda', generator=g)

    for sizes, seed in (([3, 5, 2, 9], 1), ([3, 5, 2, 9], 2), ([4, 4, 11], 3), ([3, 5, 2, 9], 4)):
        obs, rel, sse, grp, z = batch(sizes, seed)
        with torch.no_grad():
            eager = gen(obs, rel, sse, grp, user_noise=z)
        replay = graphed(obs, rel, sse, grp, z).clone()
        assert torch.equal(replay, eager), (sizes, seed)
    assert len(graphed._graphs) == 2
