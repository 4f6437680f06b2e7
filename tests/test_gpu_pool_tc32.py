"""GPU: the fp32-grade tensor-core pooling kernel (tcgen05, fp16 hi/lo operand splits; precision 'tc32', what
precision='fp32' selects for the generator dims) against the CPU oracle, the reference goldens and the CUDA-core
kernel.  Bar (BASELINE.json north_star): pooled features within 1e-5 relative."""
import numpy as np
import pytest
import torch

from conftest import load_golden, sse_from_sizes, state_dict_of
from oracle import sgan_oracle as O

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'
TOL = 1e-5
DIMS = (16, 32, 8)


def _err(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return ((a - b).abs().max() / b.abs().max()).item()


@pytest.fixture(scope='module')
def M():
    import group_gan_gcn_gat_b200.modules as M
    from group_gan_gcn_gat_b200 import _lib
    if not _lib.lib().sgx_pool_tc32_available(*DIMS):
        pytest.fail('tc32 pooling kernel unavailable on this device / build')
    return M


def _module(M, seed, precision='tc32'):
    torch.manual_seed(seed)
    return M.PoolHiddenNet(embedding_dim=16, h_dim=32, mlp_dim=64, bottleneck_dim=8, batch_norm=False, precision=precision)


def test_fp32_precision_resolves_to_tc32_for_generator_dims(M):
    from group_gan_gcn_gat_b200 import ops
    assert M.resolve_pool_precision('fp32', 16, 32, 8) == ops.PRECISION_TC32
    assert M.resolve_pool_precision('fp32', 16, 48, 48) == ops.PRECISION_FP32
    assert M.resolve_pool_precision('fp32-simt', 16, 32, 8) == ops.PRECISION_FP32


@pytest.mark.parametrize('sizes', [[1], [8], [3, 2, 7, 13, 4, 1], [70, 2, 33], [2] * 700, [300, 64, 5], [11, 3], [128],
                                   [16] * 8, [129, 3, 200, 64, 65], [512]])
def test_pool_tc32_vs_oracle(M, sizes):
    m = _module(M, 5 + len(sizes) + sizes[0])
    sse = sse_from_sizes(sizes)
    b = int(sse[-1, 1])
    h = torch.randn(1, b, 32)
    pos = torch.rand(b, 2) * 15
    ref = O.pool_hidden_net(h, sse, pos, m.state_dict())
    m = m.to(DEV)
    out = m(h.to(DEV), sse.to(DEV), pos.to(DEV))
    assert _err(out, ref) < TOL, 'tc32 vs oracle: %.3e' % _err(out, ref)


@pytest.mark.parametrize('name', ['pool_g', 'pool_g_big'])
def test_pool_tc32_fwd_bwd_vs_golden(M, name):
    """forward on the tensor cores, backward through the argmax-sparse fp32 kernels: every gradient of the
    reference's autograd (the argmax the tensor-core forward reports drives the backward)."""
    g = load_golden(name)
    m = _module(M, 0)
    m.load_state_dict(state_dict_of(g), strict=True)
    m = m.to(DEV)
    h = g['h'].to(DEV).requires_grad_(True)
    pos = g['pos'].to(DEV).requires_grad_(True)
    out = m(h, g['seq_start_end'].to(DEV), pos)
    assert _err(out, g['out']) < TOL
    (out * g['upstream'].to(DEV)).sum().backward()
    assert _err(h.grad, g['grad_in.h']) < 2e-5
    assert _err(pos.grad, g['grad_in.pos']) < 2e-5
    floor = 1e-2 * max(float(v.abs().max()) for k, v in g.items() if k.startswith('grad.'))
    for k, p in m.named_parameters():
        ref = g['grad.' + k].double()
        e = (p.grad.double().cpu() - ref).abs().max() / max(floor, ref.abs().max())
        assert e < 2e-5, (k, float(e))


def test_pool_tc32_dense_crowd_matches_cuda_core_kernel(M):
    """N = 1024: 8192 tiles over 148 persistent CTAs (the TMEM ring wraps many times); vs the fp32 CUDA-core kernel."""
    m = _module(M, 2).to(DEV)
    n = 1024
    torch.manual_seed(3)
    h = torch.randn(n, 32, device=DEV)
    pos = torch.rand(n, 2, device=DEV) * 15
    sse = sse_from_sizes([n]).to(DEV)
    a = m(h, sse, pos)
    m.precision = 'fp32-simt'
    b = m(h, sse, pos)
    assert _err(a, b) < TOL, '%.3e' % _err(a, b)
    # run-to-run determinism of the tensor-core path (max is order independent)
    m.precision = 'tc32'
    assert torch.equal(m(h, sse, pos), a)


def test_pool_tc32_argmax_matches_oracle(M):
    from group_gan_gcn_gat_b200 import ops
    from group_gan_gcn_gat_b200.schedule import get_schedule
    g = load_golden('pool_g')
    sd = state_dict_of(g)
    sched = get_schedule(g['seq_start_end'], DEV)
    out, arg = ops.pool_fwd(g['h'].reshape(-1, 32).to(DEV), g['pos'].to(DEV), sched.ped_start, sched.ped_end,
                            sched.pair_off, sched.tile_first, sched.n_pairs,
                            *[sd[k].to(DEV) for k in ('spatial_embedding.weight', 'spatial_embedding.bias',
                                                      'mlp_pre_pool.0.weight', 'mlp_pre_pool.0.bias',
                                                      'mlp_pre_pool.2.weight', 'mlp_pre_pool.2.bias')],
                            ops.PRECISION_TC32)
    ref_v, ref_i = O.pool_hidden_net_argmax(g['h'], g['seq_start_end'], g['pos'], sd)
    arg = arg.cpu()
    clear = ref_v > 1e-4                      # the winner of a near-tie may legitimately differ
    assert (arg[clear].long() == ref_i[clear]).float().mean() > 0.98
    for (s, e) in O.scene_bounds(g['seq_start_end']):
        assert ((arg[s:e] >= s) & (arg[s:e] < e)).all()


@pytest.mark.parametrize('pos_scale,h_scale', [(1e4, 1.0), (15.0, 3e4), (1e-3, 1e-3)])
def test_pool_tc32_operand_scaling_keeps_fp16_in_range(M, pos_scale, h_scale):
    """Coordinates in the 10^4 range (pixels / UTM offsets inside a scene) or hidden states of 3e4 would overflow fp16
    operands; the per-call power-of-two scale keeps the kernel finite and within the contract relative to the output
    scale.  Tiny inputs exercise fp16 subnormal lo parts."""
    m = _module(M, 11)
    sizes = [40, 3, 17]
    sse = sse_from_sizes(sizes)
    b = sum(sizes)
    torch.manual_seed(4)
    h = torch.randn(1, b, 32) * h_scale
    pos = torch.rand(b, 2) * pos_scale
    ref = O.pool_hidden_net(h.double(), sse, pos.double(), {k: v.double() for k, v in m.state_dict().items()})
    m = m.to(DEV)
    out = m(h.to(DEV), sse.to(DEV), pos.to(DEV))
    assert bool(torch.isfinite(out).all())
    assert _err(out, ref) < 2e-5, '%.3e' % _err(out, ref)


def test_pool_tc32_nan_propagates(M):
    m = _module(M, 12).to(DEV)
    sse = sse_from_sizes([5, 4]).to(DEV)
    torch.manual_seed(1)
    h = torch.randn(9, 32, device=DEV)
    pos = torch.rand(9, 2, device=DEV)
    h[6, 3] = float('nan')
    for precision in ('tc32', 'fp32-simt', 'bf16'):
        m.precision = precision
        out = m(h, sse, pos)
        assert bool(torch.isfinite(out[:5]).all()), (precision, out)     # other scene untouched
        assert bool(torch.isnan(out[5:]).all()), (precision, out)        # torch.max propagates the NaN of pair (i, 6)


def test_prepared_weights_follow_parameter_updates(M):
    """The operand images are cached per weight version: an in-place update (optimizer step, load_state_dict) must
    invalidate them."""
    m = _module(M, 13).to(DEV)
    sse = sse_from_sizes([6, 9]).to(DEV)
    torch.manual_seed(2)
    h = torch.randn(15, 32, device=DEV)
    pos = torch.rand(15, 2, device=DEV) * 10
    for precision in ('tc32', 'bf16', 'fp32-simt'):
        m.precision = precision
        a = m(h, sse, pos)
        assert torch.equal(m(h, sse, pos), a)
        with torch.no_grad():
            m.mlp_pre_pool[2].bias.add_(0.25)
        b = m(h, sse, pos)
        ref = O.pool_hidden_net(h.cpu(), sse.cpu(), pos.cpu(), {k: v.cpu() for k, v in m.state_dict().items()})
        assert _err(b, ref) < (2e-2 if precision == 'bf16' else TOL), precision
        assert not torch.equal(a, b)


def test_tc32_unsupported_dims_raise(M):
    m = M.PoolHiddenNet(embedding_dim=16, h_dim=48, mlp_dim=64, bottleneck_dim=48, batch_norm=False, precision='tc32').to(DEV)
    with pytest.raises(NotImplementedError):
        m(torch.randn(5, 48, device=DEV), torch.tensor([[0, 5]]), torch.rand(5, 2, device=DEV))
