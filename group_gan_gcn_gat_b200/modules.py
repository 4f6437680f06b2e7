"""nn.Module mirror of the reference's social-interaction modules, backed by libsgx_b200.so.

Same constructor arguments, forward signatures and ``state_dict`` key names as sgan/models.py, so
``checkpoint['g_state']`` / ``['d_state']`` load with ``strict=True`` and the reference's callers
(sgan/models.py:164, 880, 905, 902, 989) work unchanged:

    PoolHiddenNet.forward(h_states, seq_start_end, end_pos)             sgan/models.py:497
    GraphAttentionLayer.forward(h, adj) / GAT.forward(x, adj)           sgan/models.py:198 / 231
    GATEncoder.forward(h_states, seq_start_end, end_pos, end_group)     sgan/models.py:254
    GCN.forward(A, X)                                                   sgan/models.py:573
    GCNModule.forward(h_states, seq_start_end, end_pos, end_group)      sgan/models.py:628

The per-scene python loops of the reference are gone: every forward is a handful of kernel launches
over the whole ragged batch.  Unsupported reference options raise instead of silently changing
numerics: batch_norm=1 / dropout>0 inside the fused ops (never used by any shipped checkpoint).
"""
import os

import torch
import torch.nn as nn

from . import _lib, ops
from .schedule import get_schedule


def make_mlp(dim_list, activation='relu', batch_norm=True, dropout=0):
    """[Linear, (BatchNorm1d), ReLU|LeakyReLU, (Dropout)] per consecutive dim pair -- the activation also
    follows the LAST Linear, as in sgan/models.py:7-20 (pooled features and D scores are >= 0)."""
    mods = []
    for d_in, d_out in zip(dim_list, dim_list[1:]):
        mods.append(nn.Linear(d_in, d_out))
        if batch_norm:
            mods.append(nn.BatchNorm1d(d_out))
        if activation == 'relu':
            mods.append(nn.ReLU())
        elif activation == 'leakyrelu':
            mods.append(nn.LeakyReLU())
        if dropout > 0:
            mods.append(nn.Dropout(p=dropout))
    return nn.Sequential(*mods)


POOL_PRECISIONS = ('fp32', 'fp32-simt', 'tc32', 'bf16')


def resolve_pool_precision(name, embedding_dim, h_dim, bottleneck_dim):
    """Precision name -> code of include/sgx.h.

    'fp32' (the default, 1e-5 contract): the tcgen05 kernel with fp16 hi/lo operand splits ('tc32') where it is built
    for the dims -- (h_dim, bottleneck_dim) = (32, 8), the generator's pool_net of every shipped checkpoint -- and
    the CUDA-core kernel otherwise; 'fp32-simt' forces the CUDA-core kernel, 'tc32' the tensor-core one (raises for
    other dims); 'bf16' (2e-2 pooled features) is the bf16-operand tensor-core kernel.
    """
    name = (name or 'fp32').lower()
    if name in ('fp32', 'float32'):
        from . import _lib
        ok = _lib.lib().sgx_pool_tc32_available(int(embedding_dim), int(h_dim), int(bottleneck_dim))
        return ops.PRECISION_TC32 if ok else ops.PRECISION_FP32
    if name in ('fp32-simt', 'fp32_simt', 'simt'):
        return ops.PRECISION_FP32
    if name in ('tc32', 'fp32-tc', 'split'):
        return ops.PRECISION_TC32
    if name in ('bf16', 'bfloat16'):
        return ops.PRECISION_BF16
    raise ValueError('unknown pooling precision %r (%s)' % (name, ' | '.join(POOL_PRECISIONS)))


class PoolHiddenNet(nn.Module):
    """Pairwise social pooling (sgan/models.py:458-549), fused: no N^2 x 512 tensor is materialised.

    ``precision``: 'fp32' (default; 1e-5 parity -- tensor cores with fp16 hi/lo operand splits for the generator
    dims, CUDA cores otherwise), 'fp32-simt', 'tc32' or 'bf16' (tcgen05 with bf16 operands, 2e-2); see
    ``resolve_pool_precision``.  Can also be set process-wide with SGX_POOL_PRECISION.
    """

    def __init__(self, embedding_dim=64, h_dim=64, mlp_dim=1024, bottleneck_dim=1024, activation='relu',
                 batch_norm=True, dropout=0.0, precision=None):
        super().__init__()
        self.mlp_dim = 1024                      # kept (and ignored) like the reference, models.py:467
        self.h_dim = h_dim
        self.bottleneck_dim = bottleneck_dim
        self.embedding_dim = embedding_dim
        self.activation = activation
        self.batch_norm = bool(batch_norm)
        self.dropout = float(dropout)
        self.precision = precision or os.environ.get('SGX_POOL_PRECISION', 'fp32')
        self.spatial_embedding = nn.Linear(2, embedding_dim)
        self.mlp_pre_pool = make_mlp([embedding_dim + h_dim, 512, bottleneck_dim], activation=activation,
                                     batch_norm=batch_norm, dropout=dropout)

    def _fused_params(self):
        if self.batch_norm or self.activation != 'relu' or (self.dropout > 0 and self.training):
            raise NotImplementedError(
                'PoolHiddenNet fused kernel supports activation=relu, batch_norm=0, dropout=0 (every shipped '
                'checkpoint); got activation=%s batch_norm=%s dropout=%s' % (self.activation, self.batch_norm, self.dropout))
        linears = [m for m in self.mlp_pre_pool if isinstance(m, nn.Linear)]
        return linears[0], linears[1]

    def forward(self, h_states, seq_start_end, end_pos):
        l1, l2 = self._fused_params()
        sched = get_schedule(seq_start_end, end_pos.device)
        h = h_states.reshape(-1, self.h_dim)
        if h.shape[0] != sched.batch:
            raise ValueError('h_states has %d rows but seq_start_end covers %d pedestrians' % (h.shape[0], sched.batch))
        params = (self.spatial_embedding.weight, self.spatial_embedding.bias, l1.weight, l1.bias, l2.weight, l2.bias)
        code = resolve_pool_precision(self.precision, self.embedding_dim, self.h_dim, self.bottleneck_dim)
        out, _ = ops.call(ops.pool_fwd, h, end_pos, sched.ped_start, sched.ped_end, sched.pair_off, sched.tile_first,
                          sched.n_pairs, *params, code, self._prepared(params, code))
        return out

    def _prepared(self, params, code):
        """Folded first layer + tensor-core operand images, rebuilt only when a parameter changed (version counter /
        storage) -- three small launches per call otherwise, 13 calls per forward with per-step pooling."""
        key = (code,) + tuple((p.data_ptr(), p._version, p.device) for p in params)
        cached = getattr(self, '_prep_cache', None)
        if cached is not None and cached[0] == key:
            return cached[1]
        with torch.no_grad():
            buf = ops.pool_prep(*params, code, out=None if cached is None else cached[1])
        object.__setattr__(self, '_prep_cache', (key, buf))
        return buf


def _needs_grad(*tensors):
    return torch.is_grad_enabled() and any(t.requires_grad for t in tensors)


def _groups_for(sched, end_group):
    return ops.call(ops.group_ids, end_group.reshape(-1).float(), sched.ped_start, sched.ped_end, sched.scene_start)


class GraphAttentionLayer(nn.Module):
    """Dense-adjacency GAT layer (sgan/models.py:184-220).  Parameters ``W`` [in,out], ``a`` [2*out,1]."""

    def __init__(self, in_features, out_features, dropout, alpha, concat=True):
        super().__init__()
        self.dropout = dropout
        self.in_features = in_features
        self.out_features = out_features
        self.alpha = alpha
        self.concat = concat
        self.W = nn.Parameter(torch.empty(in_features, out_features))
        nn.init.xavier_uniform_(self.W.data, gain=1.414)
        self.a = nn.Parameter(torch.empty(2 * out_features, 1))
        nn.init.xavier_uniform_(self.a.data, gain=1.414)

    def forward(self, h, adj):
        if self.dropout > 0 and self.training:
            raise NotImplementedError('attention dropout > 0 is not supported by the sgx kernels')
        from . import dense
        return dense.gat_layer(h, adj, self.W, self.a, float(self.alpha), bool(self.concat))


class GAT(nn.Module):
    """n_heads concat layers + out_att, ELU, log_softmax over features (sgan/models.py:222-237)."""

    def __init__(self, nfeat, nhid, nclass, dropout, alpha, nheads):
        super().__init__()
        self.dropout = dropout
        self.nheads = nheads
        self.alpha = alpha
        for k in range(nheads):
            self.add_module('attention_%d' % k, GraphAttentionLayer(nfeat, nhid, dropout=dropout, alpha=alpha, concat=True))
        self.out_att = GraphAttentionLayer(nhid * nheads, nclass, dropout=dropout, alpha=alpha, concat=False)

    @property
    def attentions(self):
        return [getattr(self, 'attention_%d' % k) for k in range(self.nheads)]

    def forward(self, x, adj):
        if self.dropout > 0 and self.training:
            raise NotImplementedError('GAT dropout > 0 is not supported by the sgx kernels')
        x = torch.cat([att(x, adj) for att in self.attentions], dim=1)
        x = nn.functional.elu(self.out_att(x, adj))
        return nn.functional.log_softmax(x, dim=1)

    def stacked(self):
        """-> (W [heads,in,hid], a [heads,2*hid], Wout [heads*hid,out], aout [2*out]) for the fused encoder op."""
        atts = self.attentions
        if len(atts) == 1:                                  # views: no stacking kernels on the single-head path
            return atts[0].W.unsqueeze(0), atts[0].a.reshape(1, -1), self.out_att.W, self.out_att.a.reshape(-1)
        return (torch.stack([l.W for l in atts], 0), torch.stack([l.a.reshape(-1) for l in atts], 0),
                self.out_att.W, self.out_att.a.reshape(-1))


class GATEncoder(nn.Module):
    """Group-aware graph attention context (sgan/models.py:239-294), one fused op over the ragged batch."""

    def __init__(self, n_units, n_heads, dropout, alpha):
        super().__init__()
        self.n_heads = n_heads
        self.alpha = alpha
        self.dropout = dropout
        self.gat_intra = GAT(40, 72, 16, dropout, alpha, n_heads)    # dims hard-coded upstream, models.py:242-244
        self.gat_inter = GAT(16, 72, 16, dropout, alpha, n_heads)
        self.out_embedding = nn.Linear(16 * 2, 24)

    def forward(self, h_states, seq_start_end, end_pos, end_group):
        if self.dropout > 0 and self.training:
            raise NotImplementedError('GATEncoder dropout > 0 is not supported by the sgx kernels')
        sched = get_schedule(seq_start_end, h_states.device)
        if h_states.shape[0] != sched.batch:
            raise ValueError('h_states has %d rows but seq_start_end covers %d pedestrians' % (h_states.shape[0], sched.batch))
        Wi, ai, Wio, aio = self.gat_intra.stacked()
        We, ae, Weo, aeo = self.gat_inter.stacked()
        if (self.n_heads == 1 and sched.max_n <= 32 and _lib.option('graph_tc') and
                not _needs_grad(h_states, Wi, ai, Wio, aio, We, ae, Weo, aeo, self.out_embedding.weight, self.out_embedding.bias)):
            # inference: the tcgen05 kernel derives the group structure from the labels itself (no sgx_group_ids pass)
            chunk_scene, n_chunks = sched.chunks(32)
            prep = ops.gat_tc_prep(Wi, ai, Wio, aio, We, ae, Weo, aeo, self.out_embedding.weight, self.out_embedding.bias,
                                   cache=self.__dict__.setdefault('_tc_prep', {}))
            return ops.gat_encoder_fwd_labels(h_states, end_group, sched.ped_start, sched.ped_end, Wi, ai, Wio, aio, We, ae,
                                              Weo, aeo, self.out_embedding.weight, self.out_embedding.bias,
                                              float(self.alpha), sched.scene_start, chunk_scene, n_chunks, prep)
        leader, gsize, _gid, _ng = _groups_for(sched, end_group)
        # single-launch kernel: warp chunks of <= 32 peds, or <= 64 (two slots per lane) when a scene exceeds 32
        cap = 32 if sched.max_n <= 32 else 64
        chunk_scene, n_chunks = sched.chunks(cap) if self.n_heads == 1 else (sched.scene_start[:0], 0)
        return ops.call(ops.gat_encoder_fwd, h_states, leader, gsize, sched.ped_start, sched.ped_end, sched.n_scenes, Wi, ai, Wio,
                                   aio, We, ae, Weo, aeo, self.out_embedding.weight, self.out_embedding.bias,
                                   float(self.alpha), sched.scene_start, chunk_scene, n_chunks, cap, int(sched.max_n))


class GCN(nn.Module):
    """H <- ReLU((A H) W_l) for gcn_layers layers; parameters ``W.0 .. W.{L-1}`` (sgan/models.py:552-580)."""

    def __init__(self, input_dim=48, hidden_dim=72, out_dim=8, gcn_layers=2):
        super().__init__()
        self.X_dim = input_dim
        self.hidden_dim = hidden_dim
        self.out_dim = out_dim
        self.gcn_layers = gcn_layers
        self.W = nn.ParameterList()
        for l in range(gcn_layers):
            d_in = input_dim if l == 0 else hidden_dim
            d_out = out_dim if (l == gcn_layers - 1 and l > 0) else hidden_dim
            self.W.append(nn.Parameter(torch.randn(d_in, d_out)))

    def forward(self, A, X):
        from . import dense
        h = X
        for l in range(self.gcn_layers):
            h = dense.gcn_layer(A, h, self.W[l])
        return h


class GCNModule(nn.Module):
    """Group-aware graph convolution context (sgan/models.py:583-712), fused segmented-mean kernels."""

    def __init__(self, input_dim=40, hidden_dim=72, out_dim=16, gcn_layers=2, final_dim=24):
        super().__init__()
        self.gcn_layers = gcn_layers
        self.gcn_intra = GCN(input_dim=input_dim, hidden_dim=hidden_dim, out_dim=out_dim, gcn_layers=gcn_layers)
        self.gcn_inter = GCN(input_dim=16, hidden_dim=hidden_dim, out_dim=out_dim, gcn_layers=gcn_layers)
        self.out_embedding = nn.Linear(out_dim * 2, final_dim)

    @staticmethod
    def _chunks(sched):
        # Scenes of <= 32 peds (every ETH/UCY split but univ) take the single-launch warp-per-chunk kernel with the linear
        # maps on the tensor cores (0.13 ms at 245 k peds against 0.22 ms for the three-kernel path, which stays the
        # general path and what SGX_GCN_FUSED=0 selects).
        if os.environ.get('SGX_GCN_FUSED', '1') != '0':
            return sched.chunks(32)
        return sched.scene_start[:0], 0

    def forward(self, h_states, seq_start_end, end_pos, end_group):
        if self.gcn_layers != 2:
            raise NotImplementedError('GCNModule fused kernel is built for gcn_layers=2 (the reference wiring)')
        sched = get_schedule(seq_start_end, h_states.device)
        if h_states.shape[0] != sched.batch:
            raise ValueError('h_states has %d rows but seq_start_end covers %d pedestrians' % (h_states.shape[0], sched.batch))
        chunk_scene, n_chunks = self._chunks(sched)
        ws = (self.gcn_intra.W[0], self.gcn_intra.W[1], self.gcn_inter.W[0], self.gcn_inter.W[1], self.out_embedding.weight,
              self.out_embedding.bias)
        if (n_chunks > 0 and _lib.option('graph_tc') and ws[0].shape[1] == 72 and ws[1].shape[1] == 16 and
                ws[0].shape[0] in (32, 40) and ws[4].shape[0] in (24, 32) and not _needs_grad(h_states, *ws)):
            # inference: the tcgen05 kernel derives the group structure from the labels itself (no sgx_group_ids pass)
            prep = ops.gcn_tc_prep(*ws, cache=self.__dict__.setdefault('_tc_prep', {}))
            return ops.gcn_module_fwd_labels(h_states, end_group, sched.ped_start, sched.ped_end, sched.scene_start, *ws,
                                             chunk_scene, n_chunks, prep)
        leader, gsize, _gid, ngrp = _groups_for(sched, end_group)
        return ops.call(ops.gcn_module_fwd, h_states, leader, gsize, sched.ped_start, sched.ped_end, sched.scene_start, ngrp,
                                  self.gcn_intra.W[0], self.gcn_intra.W[1], self.gcn_inter.W[0], self.gcn_inter.W[1],
                                  self.out_embedding.weight, self.out_embedding.bias, chunk_scene, n_chunks)
