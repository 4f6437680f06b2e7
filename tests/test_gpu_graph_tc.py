"""GPU parity of the tcgen05 GATEncoder / GCNModule forwards (csrc/sgx_gat_tc.cu, csrc/sgx_gcn_tc.cu): against the CPU
oracle (fp32 and fp64 evaluation of sgan/models.py:254-294, 583-712), against the mma.sync kernels they replace
(library switch graph_tc = 0), and the in-kernel group structure (labels entry points) against sgx_group_ids.
Contract: 1e-5 of the largest output."""
import numpy as np
import pytest
import torch

from conftest import sse_from_sizes
from oracle import sgan_oracle as O

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


def close(a, b, tol, what):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    assert a.shape == b.shape
    e = ((a - b).abs().max() / max(1e-6, b.abs().max())).item()
    assert e <= tol, '%s: relative error %.3e > %.1e' % (what, e, tol)


def modules(in_dim=40, final=24):
    import group_gan_gcn_gat_b200.modules as M
    gat = M.GATEncoder(n_units=None, n_heads=1, dropout=0, alpha=0.2)
    gcn = M.GCNModule(input_dim=in_dim, hidden_dim=72, out_dim=16, gcn_layers=2, final_dim=final)
    with torch.no_grad():
        for p in gcn.parameters():
            if p.dim() == 2 and p.shape[0] != final:
                p.mul_(0.15)                                   # plain randn GCN weights: keep activations O(1)
    return gat, gcn


def with_option(name, val, fn):
    from group_gan_gcn_gat_b200 import _lib
    _lib.set_option(name, val)
    try:
        return fn()
    finally:
        _lib.set_option(name, 1)


def batch_of(sizes, seed, in_dim=40, label_hi=5):
    rng = np.random.RandomState(seed)
    torch.manual_seed(seed)
    n = sum(sizes)
    labs = torch.tensor(np.where(rng.rand(n) < 0.2, 0, rng.randint(1, label_hi, size=n)), dtype=torch.float32).view(-1, 1)
    return sse_from_sizes(sizes), torch.randn(n, in_dim), torch.rand(n, 2), labs


LAYOUTS = [[1], [32], [32, 32, 1], [31, 2, 32, 1, 1, 30], [1] * 70, [3, 5, 2] * 40, [7] * 18 + [32] * 3 + [1] * 5,
           list(range(1, 33)) * 2]


@pytest.mark.parametrize('sizes', LAYOUTS)
def test_tc_forwards_vs_oracle_and_mma(sizes):
    """every tile shape: a lone pedestrian, full chunks, partial tiles (fewer than four chunks), several tiles per group"""
    sse, x, pos, labs = batch_of(sizes, sum(sizes) + len(sizes))
    gat, gcn = modules()
    ref_gat = O.gat_encoder(x, sse, pos, labs, gat.state_dict(), '', 0.2, 1)
    ref_gcn = O.gcn_module(x, sse, pos, labs, gcn.state_dict(), '')
    args = (x.to(DEV), sse.to(DEV), pos.to(DEV), labs.to(DEV))
    for name, mod, ref in (('gat', gat.to(DEV), ref_gat), ('gcn', gcn.to(DEV), ref_gcn)):
        with torch.no_grad():
            out_labels = mod(*args)                                       # inference: group structure inside the kernel
            out_mma = with_option('graph_tc', 0, lambda: mod(*args))
        xg = args[0].clone().requires_grad_(True)
        out_arrays = mod(xg, *args[1:])                                   # autograd: leader / size arrays from sgx_group_ids
        close(out_labels, ref, 1e-5, '%s tcgen05 (labels) vs oracle' % name)
        close(out_mma, ref, 1e-5, '%s mma.sync vs oracle' % name)
        close(out_labels, out_mma, 5e-6, '%s tcgen05 vs mma.sync' % name)
        assert torch.equal(out_labels, out_arrays.detach()), '%s: in-kernel group structure differs from sgx_group_ids' % name


@pytest.mark.parametrize('alpha', [0.0, 0.05, 1.0, -0.3])
def test_gat_leaky_relu_slopes(alpha):
    """The one-read softmax max pass of the tcgen05 kernel relies on lrelu being monotone (alpha >= 0) and falls back to
    the per-neighbour form for a negative slope; forward (tcgen05 and mma.sync) against the oracle, and the single-launch
    backward (aggregated inter layer, d(We) / d(ae) folded in the reduction) against autograd through the oracle."""
    import group_gan_gcn_gat_b200.modules as M
    sizes = [9, 32, 1, 17, 5, 28, 3]
    sse, x, pos, labs = batch_of(sizes, 77)
    gat = M.GATEncoder(n_units=None, n_heads=1, dropout=0, alpha=alpha)
    sd = {k: v.detach().clone().requires_grad_(True) for k, v in gat.state_dict().items()}
    xr = x.clone().requires_grad_(True)
    ref = O.gat_encoder(xr, sse, pos, labs, sd, '', alpha, 1)
    up = torch.randn_like(ref)
    (ref * up).sum().backward()
    gat = gat.to(DEV)
    args = (sse.to(DEV), pos.to(DEV), labs.to(DEV))
    with torch.no_grad():
        close(gat(x.to(DEV), *args), ref, 1e-5, 'tcgen05 forward, alpha %g' % alpha)
        close(with_option('graph_tc', 0, lambda: gat(x.to(DEV), *args)), ref, 1e-5, 'mma.sync forward, alpha %g' % alpha)
    xg = x.to(DEV).requires_grad_(True)
    (gat(xg, *args) * up.to(DEV)).sum().backward()
    close(xg.grad, xr.grad, 2e-5, 'd(x), alpha %g' % alpha)
    floor = max(float(v.grad.abs().max()) for v in sd.values())
    for name, p_ in gat.named_parameters():
        g_ref = sd[name].grad
        err = float((p_.grad.cpu().double() - g_ref.double()).abs().max())
        assert err <= 1e-4 * max(floor, 1e-6), 'd(%s), alpha %g: %.3e (scale %.3e)' % (name, alpha, err, floor)


@pytest.mark.parametrize('in_dim,final', [(32, 24), (32, 32), (40, 32)])
def test_gcn_tc_every_built_instance(in_dim, final):
    sizes = [4, 32, 9, 1, 1, 17, 30, 2, 2, 8]
    sse, x, pos, labs = batch_of(sizes, 17 + in_dim + final, in_dim)
    _gat, gcn = modules(in_dim, final)
    ref = O.gcn_module(x, sse, pos, labs, gcn.state_dict(), '')
    with torch.no_grad():
        out = gcn.to(DEV)(x.to(DEV), sse.to(DEV), pos.to(DEV), labs.to(DEV))
    close(out, ref, 1e-5, 'gcn %d/%d' % (in_dim, final))


def test_in_kernel_group_structure_label_edge_cases():
    """float == semantics of the reference (sgan/models.py:263-267): +0.0 / -0.0 / NaN labels stand alone, fractional and
    negative labels group by value, equal labels in DIFFERENT scenes of one chunk stay apart."""
    sizes = [6, 6, 5, 7, 8]
    sse, x, pos, _ = batch_of(sizes, 3)
    nan = float('nan')
    labs = torch.tensor([1, 1, 0.0, -0.0, 2.5, 2.5,
                         1, 1, 2.5, nan, nan, -3,
                         -3, -3, 0, 1, 1e-30,
                         1e-30, 7, 7, 7, 7, 7, 7,
                         5, 0, 5, 0, 5, -0.0, 5, nan], dtype=torch.float32).view(-1, 1)
    gat, gcn = modules()
    from group_gan_gcn_gat_b200.schedule import get_schedule
    from group_gan_gcn_gat_b200 import ops
    sched = get_schedule(sse.to(DEV), torch.device(DEV))
    leader, gsize, _gid, _ng = ops.group_ids(labs.to(DEV), sched.ped_start, sched.ped_end, sched.scene_start)
    assert gsize.tolist() == [2, 2, 1, 1, 2, 2, 2, 2, 1, 1, 1, 1, 2, 2, 1, 1, 1, 1, 6, 6, 6, 6, 6, 6, 4, 1, 4, 1, 4, 1, 4, 1]
    args = (x.to(DEV), sse.to(DEV), pos.to(DEV), labs.to(DEV))
    for name, mod in (('gat', gat.to(DEV)), ('gcn', gcn.to(DEV))):
        with torch.no_grad():
            out_labels = mod(*args)
        out_arrays = mod(args[0].clone().requires_grad_(True), *args[1:]).detach()
        assert bool(torch.isfinite(out_labels).all())
        assert torch.equal(out_labels, out_arrays), name


@pytest.mark.parametrize('scale', [1e-6, 1e-3, 30.0, 1e3, 1e5])
def test_tc_operand_scaling_extreme_magnitudes(scale):
    """fp16 operand splits need the per-row power-of-two scaling outside [2^-2, 2^15): tiny and huge activations against an
    fp64 evaluation of the reference (softmax saturation at large scales amplifies any rounding of the scores, so the bar
    is the error of the fp32 reference itself, with a floor at the 1e-5 contract)."""
    sizes = [3, 32, 7, 1, 19, 12, 28, 4]
    sse, x, pos, labs = batch_of(sizes, 9, label_hi=4)
    x = x * scale
    gat, gcn = modules()
    d = lambda sd: {k: v.double() for k, v in sd.items()}
    ref_gat64 = O.gat_encoder(x.double(), sse, pos.double(), labs.double(), d(gat.state_dict()), '', 0.2, 1)
    ref_gcn64 = O.gcn_module(x.double(), sse, pos.double(), labs.double(), d(gcn.state_dict()), '')
    ref_gat32 = O.gat_encoder(x, sse, pos, labs, gat.state_dict(), '', 0.2, 1)
    ref_gcn32 = O.gcn_module(x, sse, pos, labs, gcn.state_dict(), '')
    args = (x.to(DEV), sse.to(DEV), pos.to(DEV), labs.to(DEV))
    for name, mod, r64, r32 in (('gat', gat.to(DEV), ref_gat64, ref_gat32), ('gcn', gcn.to(DEV), ref_gcn64, ref_gcn32)):
        with torch.no_grad():
            out = mod(*args)
        assert bool(torch.isfinite(out).all())
        ref_err = ((r32.double() - r64).abs().max() / r64.abs().max()).item()
        close(out, r64, max(1e-5, 4 * ref_err), '%s at scale %g (fp32 reference itself: %.2e)' % (name, scale, ref_err))


def test_tc_nan_and_inf_rows_stay_local():
    """a NaN / inf input row poisons its own scene only (rows of other scenes share the MMA tile but not the result)"""
    sizes = [4, 5, 3, 6]
    sse, x, pos, labs = batch_of(sizes, 11)
    x[5, 3] = float('nan')          # scene 1
    x[13, 0] = float('inf')         # scene 3
    gat, gcn = modules()
    args = (x.to(DEV), sse.to(DEV), pos.to(DEV), labs.to(DEV))
    clean = [0, 1, 2, 3, 9, 10, 11]
    xc = x.clone()
    xc[5, 3] = 0.0
    xc[13, 0] = 0.0
    for name, mod in (('gat', gat.to(DEV)), ('gcn', gcn.to(DEV))):
        with torch.no_grad():
            out = mod(*args)
            ref = mod(xc.to(DEV), *args[1:])
        assert bool(torch.isfinite(out[clean]).all()), name
        # (not bit-equal: a non-finite row maximum sends its WARP through the scaled operand path, whose fp16 split of
        # the other rows rounds differently in the last bits)
        close(out[clean], ref[clean], 2e-6, name)
        assert not bool(torch.isfinite(out[4:9]).all()), name          # the reference propagates the NaN through the scene


# ------------------------------------------------------------------ pooling backward: scene-owned kernel vs atomic kernel
def _pool_bwd_both(sizes, dims, seed, need_pos):
    """(new scene-owned path through the op, old atomic path through sgx_pool_bwd) on the same forward"""
    import ctypes  # noqa: F401
    import group_gan_gcn_gat_b200.modules as M
    from group_gan_gcn_gat_b200 import _lib, ops
    from group_gan_gcn_gat_b200.schedule import get_schedule
    E, H, B = dims
    torch.manual_seed(seed)
    sse = sse_from_sizes(sizes).to(DEV)
    n = sum(sizes)
    pool = M.PoolHiddenNet(embedding_dim=E, h_dim=H, mlp_dim=64, bottleneck_dim=B, batch_norm=False).to(DEV)
    h = torch.randn(n, H, device=DEV, requires_grad=True)
    pos = (torch.rand(n, 2, device=DEV) * 10).requires_grad_(need_pos)
    out = pool(h, sse, pos)
    go = torch.randn_like(out)
    params = list(pool.parameters())
    new = torch.autograd.grad(out, [h] + ([pos] if need_pos else []) + params, go)
    # the atomic kernel, straight through the C ABI
    sched = get_schedule(sse, torch.device(DEV))
    sd = {k: v.detach() for k, v in pool.named_parameters()}
    We, be = sd['spatial_embedding.weight'], sd['spatial_embedding.bias']
    W1, b1, W2, b2 = (sd['mlp_pre_pool.0.weight'], sd['mlp_pre_pool.0.bias'], sd['mlp_pre_pool.2.weight'],
                      sd['mlp_pre_pool.2.bias'])
    with torch.no_grad():
        o2, arg = ops._DIRECT[ops.pool_fwd](h.detach(), pos.detach(), sched.ped_start, sched.ped_end, sched.pair_off,
                                            sched.tile_first, sched.n_pairs, We, be, W1, b1, W2, b2,
                                            M.resolve_pool_precision(pool.precision, E, H, B), None)
    L = _lib.lib()
    g = [torch.empty_like(t) for t in (h, pos, We, be, W1, b1, W2, b2)]
    ws = torch.empty(L.sgx_pool_bwd_ws_bytes(n, E, H, B), dtype=torch.uint8, device=DEV)
    p = lambda t: t.data_ptr()
    _lib.check(L.sgx_pool_bwd(p(h.detach()), p(pos.detach()), p(o2), p(arg), p(go.contiguous()), n, p(We), p(be), p(W1), p(b1),
                              p(W2), p(b2), E, H, B, *[p(t) for t in g], p(ws), ws.numel(),
                              torch.cuda.current_stream().cuda_stream), 'sgx_pool_bwd')
    torch.cuda.synchronize()
    old = [g[0]] + ([g[1]] if need_pos else []) + [g[2], g[3], g[4], g[5], g[6], g[7]]
    return new, old


@pytest.mark.parametrize('sizes', [[1], [2, 3], [16] * 3, [33, 1, 15], [14, 2, 34, 3, 9], [1] * 40 + [33], [64, 5, 5], [7] * 30,
                                   [48], [49, 1]])
@pytest.mark.parametrize('dims', [(16, 32, 8), (16, 48, 48)])
@pytest.mark.parametrize('need_pos', [False, True])
def test_pool_backward_scene_kernel_matches_atomic_kernel(sizes, dims, need_pos):
    """blocks of whole scenes (<= 48 rows in shared memory), blocks with no scene start, scenes too large for the tile
    (atomic fall-back inside the kernel), generator and discriminator dims, with and without the position gradient"""
    new, old = _pool_bwd_both(sizes, dims, sum(sizes) + dims[2], need_pos)
    scale = max(float(t.abs().max()) for t in old)
    for i, (a, b) in enumerate(zip(new, old)):
        assert a.shape == b.shape
        err = float((a - b).abs().max())
        assert err <= 2e-5 * max(scale, 1e-6), 'gradient %d: %.3e vs scale %.3e' % (i, err, scale)


def test_tc_weight_images_follow_parameter_updates():
    """the cached fp16 weight images (sgx_*_tc_prep) are rebuilt after an in-place update (optimizer step, load_state_dict)
    and after a parameter is reassigned"""
    sizes = [5, 9, 2, 14]
    sse, x, pos, labs = batch_of(sizes, 21)
    gat, gcn = modules()
    args = (x.to(DEV), sse.to(DEV), pos.to(DEV), labs.to(DEV))
    for name, mod, oracle in (('gat', gat.to(DEV), lambda m: O.gat_encoder(x, sse, pos, labs, {k: v.cpu() for k, v in m.state_dict().items()}, '', 0.2, 1)),
                              ('gcn', gcn.to(DEV), lambda m: O.gcn_module(x, sse, pos, labs, {k: v.cpu() for k, v in m.state_dict().items()}, ''))):
        with torch.no_grad():
            out0 = mod(*args).clone()
            assert torch.equal(mod(*args), out0)                      # second call: cached images, same bits
            close(out0, oracle(mod), 1e-5, name)
            for p in mod.parameters():
                p.mul_(0.9)                                           # in-place update (version counter)
            out1 = mod(*args).clone()
            close(out1, oracle(mod), 1e-5, name + ' after in-place update')
            assert not torch.equal(out0, out1)
            first = next(mod.parameters())
            first.data = first.data * 1.1                             # reassigned storage (data_ptr)
            close(mod(*args), oracle(mod), 1e-5, name + ' after reassignment')


def test_autograd_function_twins_match_the_registered_ops():
    """ops.call() runs a torch.autograd.Function twin of each custom op under autograd (no torch.library dispatcher on the
    step path); the twin must give the gradients of the registered op (same forward body, setup, backward)."""
    import group_gan_gcn_gat_b200.modules as M
    from group_gan_gcn_gat_b200 import ops
    from group_gan_gcn_gat_b200.schedule import get_schedule
    sizes = [5, 12, 3, 33, 7]                      # one scene above 32: general GAT / GCN backward, fused pooling backward
    sse, x, pos, labs = batch_of(sizes, 31)
    dev = torch.device(DEV)
    sched = get_schedule(sse.to(dev), dev)
    pool = M.PoolHiddenNet(embedding_dim=16, h_dim=32, mlp_dim=64, bottleneck_dim=8, batch_norm=False).to(dev)
    h = torch.randn(sum(sizes), 32, device=dev)
    l1, l2 = pool._fused_params()
    params = (pool.spatial_embedding.weight, pool.spatial_embedding.bias, l1.weight, l1.bias, l2.weight, l2.bias)
    code = M.resolve_pool_precision(pool.precision, 16, 32, 8)
    args = (sched.ped_start, sched.ped_end, sched.pair_off, sched.tile_first, sched.n_pairs, *params, code, None)
    up = torch.randn(sum(sizes), 8, device=dev)
    grads = []
    for use_twin in (True, False):
        hh = h.clone().requires_grad_(True)
        pp = pos.to(dev).clone().requires_grad_(True)
        out, _arg = (ops.call(ops.pool_fwd, hh, pp, *args) if use_twin else ops.pool_fwd(hh, pp, *args))
        grads.append(torch.autograd.grad((out * up).sum(), [hh, pp] + list(params)))
    for i, (a, b) in enumerate(zip(*grads)):       # (the position gradient is accumulated with shared-memory atomics: not
        close(a, b, 1e-6, 'gradient %d' % i)        # bit-reproducible between two launches)
    assert ops._FAST[ops.pool_fwd] is not None and ops._FAST[ops.gat_encoder_fwd] is not None and ops._FAST[ops.gcn_module_fwd] is not None
