"""GPU: the tcgen05/TMEM bf16 pooling kernel against the CPU oracle (tolerance 2e-2, BASELINE.json north_star)
and against the fp32 CUDA-core kernel."""
import numpy as np
import pytest
import torch

from conftest import load_golden, sse_from_sizes, state_dict_of
from oracle import sgan_oracle as O

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'
TOL = 2e-2


def _err(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return ((a - b).abs().max() / b.abs().max()).item()


def _run(M, sizes, dims, seed):
    e_dim, h_dim, bott = dims
    torch.manual_seed(seed)
    m = M.PoolHiddenNet(embedding_dim=e_dim, h_dim=h_dim, mlp_dim=64, bottleneck_dim=bott, batch_norm=False)
    sse = sse_from_sizes(sizes)
    b = int(sse[-1, 1])
    h = torch.randn(1, b, h_dim)
    pos = torch.rand(b, 2) * 15
    ref = O.pool_hidden_net(h, sse, pos, m.state_dict())
    m = m.to(DEV)
    m.precision = 'fp32'
    out32 = m(h.to(DEV), sse.to(DEV), pos.to(DEV))
    m.precision = 'bf16'
    out16 = m(h.to(DEV), sse.to(DEV), pos.to(DEV))
    torch.cuda.synchronize()
    return out16, out32, ref


@pytest.fixture(scope='module')
def M():
    import group_gan_gcn_gat_b200.modules as M
    from group_gan_gcn_gat_b200 import _lib
    if not _lib.lib().sgx_has_tcgen05():
        pytest.fail('libsgx_b200.so was built without the tcgen05 pooling kernel')
    return M


@pytest.mark.parametrize('sizes', [[8], [3, 2, 7, 13, 4, 1], [70, 2, 33], [2] * 700, [300, 64, 5]])
def test_pool_bf16_generator_dims(M, sizes):
    out16, out32, ref = _run(M, sizes, (16, 32, 8), 5 + len(sizes))
    assert _err(out32, ref) < 1e-5
    assert _err(out16, ref) < TOL, 'bf16 vs oracle: %.3e' % _err(out16, ref)


@pytest.mark.parametrize('sizes', [[9], [40, 7, 130], [2] * 300])
def test_pool_bf16_discriminator_dims(M, sizes):
    out16, out32, ref = _run(M, sizes, (16, 48, 48), 9 + len(sizes))
    assert _err(out16, ref) < TOL, 'bf16 vs oracle: %.3e' % _err(out16, ref)


def test_pool_bf16_vs_golden(M):
    g = load_golden('pool_g')
    m = M.PoolHiddenNet(embedding_dim=16, h_dim=32, mlp_dim=64, bottleneck_dim=8, batch_norm=False, precision='bf16')
    m.load_state_dict(state_dict_of(g), strict=True)
    out = m.to(DEV)(g['h'].to(DEV), g['seq_start_end'].to(DEV), g['pos'].to(DEV))
    assert _err(out, g['out']) < TOL


def test_pool_bf16_dense_crowd_matches_fp32_kernel(M):
    """N = 1024 (8192 tiles, > 148 persistent CTAs x many tiles each): bf16 tensor-core kernel vs the fp32 kernel."""
    out16, out32, _ = _run(M, [64], (16, 32, 8), 1)
    torch.manual_seed(2)
    m = M.PoolHiddenNet(embedding_dim=16, h_dim=32, mlp_dim=64, bottleneck_dim=8, batch_norm=False).to(DEV)
    n = 1024
    h = torch.randn(n, 32, device=DEV)
    pos = torch.rand(n, 2, device=DEV) * 15
    sse = sse_from_sizes([n]).to(DEV)
    m.precision = 'fp32'
    a = m(h, sse, pos)
    m.precision = 'bf16'
    b = m(h, sse, pos)
    assert _err(b, a) < TOL


def test_pool_bf16_unsupported_dims_raise(M):
    m = M.PoolHiddenNet(embedding_dim=8, h_dim=20, mlp_dim=64, bottleneck_dim=24, batch_norm=False, precision='bf16').to(DEV)
    with pytest.raises(NotImplementedError):
        m(torch.randn(5, 20, device=DEV), torch.tensor([[0, 5]]), torch.rand(5, 2, device=DEV))
