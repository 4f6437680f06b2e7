"""Programmatic dependent launch along the kernels of one SGAN-P forward (encoder recurrence -> pooling statistics ->
h image -> pooling -> unpack -> context MLP -> decoder recurrence; csrc/sgx_common.cuh `launch_pdl`): the overlapped
chain must give the SAME BITS as the serialised one (option 'pdl' = 0), on a batch large enough for the tcgen05
recurrences and for several waves of every kernel, forward after forward on one stream (a kernel that touched another
kernel's buffer before its griddepcontrol.wait would show up here as a flipped bit sooner or later).
"""
import numpy as np
import pytest
import torch

from conftest import load_golden, state_dict_of

pytestmark = pytest.mark.gpu

DEV = 'cuda:0'


def _sgan_p(dev):
    import group_gan_gcn_gat_b200.models as MD
    g = load_golden('generator_p_eth')
    gen = MD.TrajectoryGenerator(obs_len=8, pred_len=8, embedding_dim=16, encoder_h_dim=32, decoder_h_dim=32, mlp_dim=64,
                                 num_layers=1, noise_dim=(8,), noise_type='gaussian', noise_mix_type='global',
                                 pooling_type='pool_net', pool_every_timestep=False, dropout=0, bottleneck_dim=8,
                                 batch_norm=False, context_type='mlp')
    gen.load_state_dict(state_dict_of(g), strict=False)
    return gen.to(dev).train()


def _batch(n_scenes, seed):
    rng = np.random.RandomState(seed)
    sizes = rng.choice([1, 2, 3, 4, 5, 7, 9, 13, 21, 40], size=n_scenes, p=[.2, .2, .15, .12, .1, .08, .06, .05, .03, .01])
    cs = np.concatenate([[0], np.cumsum(sizes)])
    n = int(cs[-1])
    sse = torch.tensor(np.stack([cs[:-1], cs[1:]], axis=1), dtype=torch.int64)
    rel = torch.from_numpy(rng.randn(8, n, 2).astype(np.float32) * 0.3)
    obs = torch.cumsum(rel, dim=0) + torch.from_numpy(rng.rand(1, n, 2).astype(np.float32) * 12)
    grp = torch.from_numpy(rng.randint(0, 3, size=(8, n, 1)).astype(np.float32))
    noise = torch.from_numpy(rng.randn(n_scenes, 8).astype(np.float32))
    return obs, rel, sse, grp, noise


@pytest.mark.parametrize('n_scenes', [6000, 30000])
def test_pdl_chain_is_bit_equal_to_the_serialised_chain(n_scenes):
    from group_gan_gcn_gat_b200 import _lib
    gen = _sgan_p(DEV)
    obs, rel, sse, grp, noise = [t.to(DEV) for t in _batch(n_scenes, 5)]
    assert obs.shape[1] >= 8192                      # the tcgen05 recurrences
    prev = _lib.option('pdl')
    try:
        _lib.set_option('pdl', 0)
        with torch.no_grad():
            ref = gen(obs, rel, sse, grp, user_noise=noise).clone()
        _lib.set_option('pdl', 1)
        with torch.no_grad():
            outs = [gen(obs, rel, sse, grp, user_noise=noise) for _ in range(12)]      # back to back, no sync between them
        torch.cuda.synchronize()
        for k, o in enumerate(outs):
            assert torch.equal(o, ref), 'forward %d of the overlapped chain differs from the serialised chain' % k
    finally:
        _lib.set_option('pdl', prev)
    assert torch.isfinite(ref).all()


def test_pdl_option_is_known_to_the_library():
    from group_gan_gcn_gat_b200 import _lib
    assert _lib.option('pdl') in (0, 1)
    _lib.set_option('pdl', _lib.option('pdl'))
