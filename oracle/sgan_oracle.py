"""CPU oracle for the sgan social-interaction path (TEST INFRASTRUCTURE ONLY).

This file is a plain-PyTorch/numpy *restatement* of the reference's algorithm for the
hot path of peaceminusones/Group-GAN-GCN-GAT.  It is the checker the CUDA path is
compared against; it is NOT part of the product.  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of
``bench.py`` may import it.  The product package never does.

Parity pin: the reference ships no tests.  This restatement is pinned
against the reference itself: ``oracle/make_golden.py`` imports ``/root/reference``
(in the build container, where it exists), runs the reference modules on seeded
inputs and freezes inputs/outputs/gradients under ``tests/golden/``;
``tests/test_oracle_golden.py`` then checks every function here against those
fixtures plus the two known-answer artefacts of the reference
(``Untitled.ipynb:546-556`` R matrix, ``sgan/GCN.py:128`` label example).

Everything is written as *functions over a state_dict* (the reference's own
parameter names), so the same weights drive the oracle and the CUDA modules.
The arithmetic follows the reference operation-for-operation (same torch ops on
the same shapes) so that CPU results agree with the reference to the last bit
where torch is deterministic.

Reference locations (relative to /root/reference):
  make_mlp                 sgan/models.py:7-20
  PoolHiddenNet.forward    sgan/models.py:497-549
  GraphAttentionLayer      sgan/models.py:184-220
  GAT.forward              sgan/models.py:231-237
  GATEncoder.forward       sgan/models.py:254-294
  GCN.forward              sgan/models.py:573-580
  GCNModule.forward        sgan/models.py:628-712
  Encoder/Decoder          sgan/models.py:32-178
  TrajectoryGenerator      sgan/models.py:862-927
  TrajectoryDiscriminator  sgan/models.py:972-991
  relative_to_abs          sgan/utils.py:83-96
  displacement errors      sgan/losses.py:74-119
"""
import numpy as np
import torch
import torch.nn.functional as F

MASK_FILL = -9e15  # sgan/models.py:202


# --------------------------------------------------------------------------------------
# ragged segmentation (sgan/data/trajectories_GCN.py:19-22, sgan/models.py:507-510)
# --------------------------------------------------------------------------------------
def scene_bounds(seq_start_end):
    """[(start, end)] python ints, in scene order."""
    sse = seq_start_end.detach().cpu().numpy() if torch.is_tensor(seq_start_end) else np.asarray(seq_start_end)
    return [(int(a), int(b)) for a, b in sse.reshape(-1, 2)]


# --------------------------------------------------------------------------------------
# group structure (integer-exact part)          sgan/models.py:263-278 / 654-680
# --------------------------------------------------------------------------------------
def group_mask(labels_col):
    """M_intra for one scene.  labels_col: float tensor [N,1].

    M[i,j] = (g_i == g_j and g_i != 0) or i == j          (sgan/models.py:263-266)
    Python precedence in the reference: ``(A == B) & (A != 0) | eye``.
    """
    n = labels_col.shape[0]
    gi = labels_col.expand(n, n)                 # row i holds g_i everywhere
    gj = labels_col.reshape(1, n).expand(n, n)   # row i holds g_0..g_{n-1}
    same = (gi == gj) & (gi != 0)
    return same | torch.eye(n, dtype=torch.bool, device=labels_col.device)


def row_normalize(mask):
    """bool matrix -> fp32 matrix with rows scaled by 1/rowsum  (sgan/models.py:246-252).

    rowsum is int64 -> .float() -> pow(-1); bool * float -> float32.
    """
    inv = mask.sum(1).unsqueeze(1).float().pow(-1)
    return mask.mul(inv)


def group_rows(mask):
    """R_intra (bool [G,N]): distinct rows of M, in the order the reference ends up with.

    torch.unique(dim=0) yields ascending lexicographic rows; the reference walks them
    last-to-first (sgan/models.py:271-277) => groups ordered by ascending minimum member.
    """
    uniq = torch.unique(mask, sorted=False, dim=0)
    return torch.flip(uniq, dims=[0])


def group_ids_numpy(labels, seq_start_end):
    """Integer oracle for the group structure, independent of torch.unique.

    labels: float array [batch]; returns dict with
      group_id   int32 [batch]  scene-local id, groups numbered by ascending min member
      group_size int32 [batch]  size of the ped's group
      leader     int32 [batch]  global index of the min member of the ped's group
      n_group    int32 [S]
    """
    labels = np.asarray(labels, dtype=np.float32).reshape(-1)
    batch = labels.shape[0]
    gid = np.zeros(batch, np.int32)
    gsz = np.zeros(batch, np.int32)
    lead = np.zeros(batch, np.int32)
    ngrp = []
    for (s, e) in scene_bounds(seq_start_end):
        nxt = 0
        seen = {}
        for i in range(s, e):
            lab = labels[i]
            if lab != 0 and float(lab) in seen:
                g, l = seen[float(lab)]
            else:
                g, l = nxt, i
                nxt += 1
                if lab != 0:
                    seen[float(lab)] = (g, l)
            gid[i] = g
            lead[i] = l
        for i in range(s, e):
            gsz[i] = int(np.sum(lead[s:e] == lead[i]))
        ngrp.append(nxt)
    return dict(group_id=gid, group_size=gsz, leader=lead, n_group=np.asarray(ngrp, np.int32))


# --------------------------------------------------------------------------------------
# make_mlp as a function over a state_dict       sgan/models.py:7-20
# --------------------------------------------------------------------------------------
def mlp_apply(x, sd, prefix, n_layers, activation='relu'):
    """Sequential [Linear, act] * n_layers with batch_norm=0, dropout=0 (indices 0,2,4..)."""
    for l in range(n_layers):
        x = F.linear(x, sd[f'{prefix}{2 * l}.weight'], sd[f'{prefix}{2 * l}.bias'])
        if activation == 'relu':
            x = F.relu(x)
        elif activation == 'leakyrelu':
            x = F.leaky_relu(x)
    return x


# --------------------------------------------------------------------------------------
# PoolHiddenNet                                   sgan/models.py:497-549
# --------------------------------------------------------------------------------------
def pool_hidden_net(h_states, seq_start_end, end_pos, sd, prefix=''):
    """out_i = max_j ReLU(W2 ReLU(W1 [We (P_j - P_i) + be ; h_j] + b1) + b2), j over i's scene.

    Materialises every ordered pair exactly as the reference does (pair p = i*N + j).
    """
    w_e, b_e = sd[prefix + 'spatial_embedding.weight'], sd[prefix + 'spatial_embedding.bias']
    h_dim = sd[prefix + 'mlp_pre_pool.0.weight'].shape[1] - w_e.shape[0]
    flat_h = h_states.reshape(-1, h_dim)
    pooled = []
    for (s, e) in scene_bounds(seq_start_end):
        n = e - s
        h = flat_h[s:e]
        p = end_pos[s:e]
        h_j = h.unsqueeze(0).expand(n, n, h_dim).reshape(n * n, h_dim)      # row p -> h_j
        p_j = p.unsqueeze(0).expand(n, n, 2).reshape(n * n, 2)              # row p -> P_j
        p_i = p.unsqueeze(1).expand(n, n, 2).reshape(n * n, 2)              # row p -> P_i
        emb = F.linear(p_j - p_i, w_e, b_e)
        z = mlp_apply(torch.cat([emb, h_j], dim=1), sd, prefix + 'mlp_pre_pool.', 2)
        pooled.append(z.view(n, n, -1).max(1)[0])
    return torch.cat(pooled, dim=0)


def pool_hidden_net_argmax(h_states, seq_start_end, end_pos, sd, prefix=''):
    """Same as pool_hidden_net but also returns the (global) neighbour index attaining the max."""
    w_e, b_e = sd[prefix + 'spatial_embedding.weight'], sd[prefix + 'spatial_embedding.bias']
    h_dim = sd[prefix + 'mlp_pre_pool.0.weight'].shape[1] - w_e.shape[0]
    flat_h = h_states.reshape(-1, h_dim)
    vals, idxs = [], []
    for (s, e) in scene_bounds(seq_start_end):
        n = e - s
        h, p = flat_h[s:e], end_pos[s:e]
        rel = (p.unsqueeze(0) - p.unsqueeze(1)).reshape(n * n, 2)
        emb = F.linear(rel, w_e, b_e)
        x = torch.cat([emb, h.unsqueeze(0).expand(n, n, h_dim).reshape(n * n, h_dim)], 1)
        z = mlp_apply(x, sd, prefix + 'mlp_pre_pool.', 2).view(n, n, -1)
        v, a = z.max(1)
        vals.append(v)
        idxs.append(a + s)
    return torch.cat(vals, 0), torch.cat(idxs, 0)


# --------------------------------------------------------------------------------------
# GraphAttentionLayer / GAT                        sgan/models.py:184-237
# --------------------------------------------------------------------------------------
def graph_attention_layer(h, adj, w, a, alpha, concat=True):
    """Dense GAT layer, dropout = 0.  Builds the [N,N,2F] pair tensor like the reference."""
    wh = torch.mm(h, w)
    n, f = wh.shape
    left = wh.repeat_interleave(n, dim=0)
    right = wh.repeat(n, 1)
    pair = torch.cat([left, right], dim=1).view(n, n, 2 * f)
    e = F.leaky_relu(torch.matmul(pair, a).squeeze(2), alpha)
    att = torch.where(adj > 0, e, MASK_FILL * torch.ones_like(e))
    att = F.softmax(att, dim=1)
    out = torch.matmul(att, wh)
    return F.elu(out) if concat else out


def gat(x, adj, sd, prefix, alpha, n_heads):
    heads = [graph_attention_layer(x, adj, sd[f'{prefix}attention_{k}.W'], sd[f'{prefix}attention_{k}.a'],
                                   alpha, True) for k in range(n_heads)]
    x = torch.cat(heads, dim=1)
    x = F.elu(graph_attention_layer(x, adj, sd[prefix + 'out_att.W'], sd[prefix + 'out_att.a'], alpha, False))
    return F.log_softmax(x, dim=1)


def _scene_graph(labels_col):
    m = group_mask(labels_col)
    a_intra = row_normalize(m)
    r = group_rows(m)
    r_n = row_normalize(r)
    g = r.shape[0]
    a_inter = row_normalize(torch.ones((g, g), dtype=torch.bool, device=labels_col.device))
    return m, a_intra, r, r_n, a_inter


def gat_encoder(h_states, seq_start_end, end_pos, end_group, sd, prefix='', alpha=0.2, n_heads=1):
    """GATEncoder.forward (sgan/models.py:254-294): intra GAT -> GPool -> inter GAT -> unpool -> Linear."""
    outs = []
    for (s, e) in scene_bounds(seq_start_end):
        x = h_states[s:e]
        _, a_intra, _, r_n, a_inter = _scene_graph(end_group[s:e])
        r_n = r_n.to(x.dtype)
        x1 = gat(x, a_intra.to(x.dtype), sd, prefix + 'gat_intra.', alpha, n_heads)
        xg = torch.matmul(r_n, x1)
        yg = gat(xg, a_inter.to(x.dtype), sd, prefix + 'gat_inter.', alpha, n_heads)
        x2 = torch.matmul(r_n.T, yg)
        outs.append(F.linear(torch.cat([x1, x2], dim=1), sd[prefix + 'out_embedding.weight'],
                             sd[prefix + 'out_embedding.bias']))
    return torch.cat(outs, dim=0)


# --------------------------------------------------------------------------------------
# GCN / GCNModule                                  sgan/models.py:552-712
# --------------------------------------------------------------------------------------
def gcn(adj, x, sd, prefix, n_layers=2):
    h = x
    for l in range(n_layers):
        h = F.relu(torch.matmul(torch.matmul(adj, h), sd[f'{prefix}W.{l}']))
    return h


def gcn_module(h_states, seq_start_end, end_pos, end_group, sd, prefix='', n_layers=2):
    outs = []
    for (s, e) in scene_bounds(seq_start_end):
        x = h_states[s:e]
        _, a_intra, _, r_n, a_inter = _scene_graph(end_group[s:e])
        r_n = r_n.to(x.dtype)
        x1 = gcn(a_intra.to(x.dtype), x, sd, prefix + 'gcn_intra.', n_layers)
        xg = torch.matmul(r_n, x1)
        yg = gcn(a_inter.to(x.dtype), xg, sd, prefix + 'gcn_inter.', n_layers)
        x2 = torch.matmul(r_n.T, yg)
        outs.append(F.linear(torch.cat([x1, x2], dim=1), sd[prefix + 'out_embedding.weight'],
                             sd[prefix + 'out_embedding.bias']))
    return torch.cat(outs, dim=0)


# --------------------------------------------------------------------------------------
# Encoder / Decoder / generator / discriminator wiring   sgan/models.py:32-178, 862-991
# --------------------------------------------------------------------------------------
def _lstm(x_seq, state, sd, prefix):
    """Single-layer LSTM through torch's own CPU kernel with the reference's parameter names."""
    flat = [sd[prefix + 'weight_ih_l0'], sd[prefix + 'weight_hh_l0'], sd[prefix + 'bias_ih_l0'], sd[prefix + 'bias_hh_l0']]
    out, h, c = torch._VF.lstm(x_seq, state, flat, True, 1, 0.0, False, False, False)
    return out, (h, c)


def encoder(obs_traj_rel, sd, prefix):
    w, b = sd[prefix + 'spatial_embedding.weight'], sd[prefix + 'spatial_embedding.bias']
    batch = obs_traj_rel.shape[1]
    h_dim = sd[prefix + 'encoder.weight_hh_l0'].shape[1]
    emb = F.linear(obs_traj_rel.reshape(-1, 2), w, b).view(-1, batch, w.shape[0])
    zeros = torch.zeros(1, batch, h_dim, dtype=emb.dtype)
    _, (h, _) = _lstm(emb, (zeros, zeros.clone()), sd, prefix + 'encoder.')
    return h


def decoder(last_pos, last_pos_rel, state, seq_start_end, sd, prefix, seq_len, pool_every_timestep):
    w, b = sd[prefix + 'spatial_embedding.weight'], sd[prefix + 'spatial_embedding.bias']
    batch = last_pos.shape[0]
    h_dim = sd[prefix + 'decoder.weight_hh_l0'].shape[1]
    x = F.linear(last_pos_rel, w, b).view(1, batch, -1)
    rels = []
    for _ in range(seq_len):
        out, state = _lstm(x, state, sd, prefix + 'decoder.')
        rel = F.linear(out.view(-1, h_dim), sd[prefix + 'hidden2pos.weight'], sd[prefix + 'hidden2pos.bias'])
        cur = rel + last_pos
        if pool_every_timestep:
            ph = pool_hidden_net(state[0], seq_start_end, cur, sd, prefix + 'pool_net.')
            hh = mlp_apply(torch.cat([state[0].view(-1, h_dim), ph], dim=1), sd, prefix + 'mlp.', 2)
            state = (hh.unsqueeze(0), state[1])
        x = F.linear(rel, w, b).view(1, batch, -1)
        rels.append(rel.view(batch, -1))
        last_pos = cur
    return torch.stack(rels, dim=0), state[0]


def add_global_noise(x, seq_start_end, z):
    """noise_mix_type='global': one noise vector per scene, repeated over its peds (models.py:837-846)."""
    parts = []
    for k, (s, e) in enumerate(scene_bounds(seq_start_end)):
        parts.append(torch.cat([x[s:e], z[k].view(1, -1).repeat(e - s, 1)], dim=1))
    return torch.cat(parts, dim=0)


def generator_forward(obs_traj, obs_traj_rel, seq_start_end, obs_traj_g, sd, cfg, user_noise):
    """TrajectoryGenerator.forward (sgan/models.py:862-927) for the three wirings of SURVEY A.2.

    cfg: dict(pred_len, wiring in {'gat','gcn','mlp'}, pooling (bool), pool_every_timestep,
              alpha, n_heads, noise_mix_type)
    """
    h = encoder(obs_traj_rel, sd, 'encoder.')
    enc_dim = h.shape[-1]
    end_pos = obs_traj[-1]
    if cfg.get('pooling', True):
        ph = pool_hidden_net(h, seq_start_end, end_pos, sd, 'pool_net.')
        ctx = torch.cat([h.view(-1, enc_dim), ph], dim=1)
    else:
        ctx = h.view(-1, enc_dim)
    grp = obs_traj_g[-1]
    if cfg['wiring'] == 'gat':
        ctx = gat_encoder(ctx, seq_start_end, end_pos, grp, sd, 'gatencoder.', cfg.get('alpha', 0.2), cfg.get('n_heads', 1))
    elif cfg['wiring'] == 'gcn':
        ctx = gcn_module(ctx, seq_start_end, end_pos, grp, sd, 'gcn_module.')
    else:
        ctx = mlp_apply(ctx, sd, 'mlp_decoder_context.', 2)
    if user_noise is not None:
        if cfg.get('noise_mix_type', 'global') == 'global':
            ctx = add_global_noise(ctx, seq_start_end, user_noise)
        else:
            ctx = torch.cat([ctx, user_noise], dim=1)
    dh = ctx.unsqueeze(0)
    dc = torch.zeros_like(dh)
    rel, _ = decoder(obs_traj[-1], obs_traj_rel[-1], (dh, dc), seq_start_end, sd, 'decoder.',
                     cfg['pred_len'], cfg.get('pool_every_timestep', False))
    return rel


def discriminator_forward(traj, traj_rel, seq_start_end, sd, d_type='global'):
    h = encoder(traj_rel, sd, 'encoder.')
    if d_type == 'local':
        x = h.squeeze()
    else:
        x = pool_hidden_net(h.squeeze(), seq_start_end, traj[0], sd, 'pool_net.')
    return mlp_apply(x, sd, 'real_classifier.', 2)


# --------------------------------------------------------------------------------------
# metrics                                         sgan/utils.py:83-96, sgan/losses.py:74-119
# --------------------------------------------------------------------------------------
def relative_to_abs(rel_traj, start_pos):
    return (torch.cumsum(rel_traj.permute(1, 0, 2), dim=1) + start_pos.unsqueeze(1)).permute(1, 0, 2)


def displacement_error_raw(pred, gt):
    d = (gt.permute(1, 0, 2) - pred.permute(1, 0, 2)) ** 2
    return torch.sqrt(d.sum(dim=2)).sum(dim=1)


def final_displacement_error_raw(pred_last, gt_last):
    return torch.sqrt(((gt_last - pred_last) ** 2).sum(dim=1))


def best_of_k(per_sample_errors, seq_start_end):
    """evaluate_helper (scripts/evaluate_model.py:58-69): per scene, sum over peds then min over K."""
    stacked = torch.stack(per_sample_errors, dim=1)
    total = 0.0
    for (s, e) in scene_bounds(seq_start_end):
        total = total + torch.min(stacked[s:e].sum(dim=0))
    return total
