"""Generator / discriminator wiring around the sgx ops -- mirror of sgan/models.py:32-178, 715-991.

Same class names, constructor keywords, forward signatures and state_dict keys as the reference, so the
reference's scripts (scripts/train.py:395-484, scripts/evaluate_model.py:72-99) can import these classes
instead.  Differences that are not visible in results:
  * no hard-coded ``.cuda()``: new tensors are created on the device of the inputs;
  * ``seq_start_end`` is consumed through a cached SceneSchedule (no per-scene ``.item()``);
  * ``context_type`` selects what the reference selects by (un)commenting lines 898-905:
      'gat' (live code, models.py:905), 'gcn' (models.py:902), 'mlp' (upstream SGAN, models.py:898).
"""
import torch
import torch.nn as nn

from . import ops
from .modules import GATEncoder, GCNModule, PoolHiddenNet, make_mlp
from .schedule import get_schedule
from .utils import ready, ready_last


def _fused_lstm_ok(module, lstm, *tensors):
    """The inference kernels (tensor-core LSTM for large batches); anything that needs autograd goes through
    _fused_lstm_train_ok -> ops.lstm_*_train, or nn.LSTM (cuDNN) as the last resort."""
    if not all(t.is_cuda and t.dtype == torch.float32 for t in tensors):
        return False
    if lstm.num_layers != 1 or lstm.hidden_size not in ops.FUSED_LSTM_H or lstm.bidirectional:
        return False
    if torch.is_grad_enabled() and (any(t.requires_grad for t in tensors) or
                                    any(p.requires_grad for p in module.parameters())):
        return False
    return True


def _fused_lstm_train_ok(module, lstm, *tensors):
    """Under autograd the recurrence runs as one forward kernel with a tape + one backward kernel
    (ops.lstm_encoder_train / lstm_decoder_train) instead of cuDNN + a python step loop.  SGX_LSTM_TRAIN=0 keeps the
    nn.LSTM path (kept for A/B parity tests)."""
    import os
    if os.environ.get('SGX_LSTM_TRAIN', '1') == '0':
        return False
    if not all(t is None or (t.is_cuda and t.dtype == torch.float32) for t in tensors):
        return False
    if lstm.num_layers != 1 or lstm.hidden_size not in ops.FUSED_LSTM_H or lstm.bidirectional:
        return False
    return not (lstm.dropout and module.training)


def get_noise(shape, noise_type, device=None):
    """Drawn on the CPU generator then moved, exactly like sgan/models.py:23-29 (seed-compatible)."""
    if noise_type == 'gaussian':
        z = torch.randn(*shape)
    elif noise_type == 'uniform':
        z = torch.rand(*shape).sub_(0.5).mul_(2.0)
    else:
        raise ValueError('Unrecognized noise type "%s"' % noise_type)
    return z.to(device) if device is not None else z


class Encoder(nn.Module):
    """Linear(2, E) + single LSTM over the observed relative displacements (sgan/models.py:32-92)."""

    def __init__(self, embedding_dim=64, h_dim=64, mlp_dim=1024, num_layers=1, dropout=0.0):
        super().__init__()
        self.mlp_dim = 1024
        self.h_dim = h_dim
        self.embedding_dim = embedding_dim
        self.num_layers = num_layers
        self.spatial_embedding = nn.Linear(2, embedding_dim)
        self.encoder = nn.LSTM(embedding_dim, h_dim, num_layers, dropout=dropout)

    def init_hidden(self, batch, like):
        z = like.new_zeros(self.num_layers, batch, self.h_dim)
        return z, z.clone()

    def forward(self, obs_traj):
        batch = obs_traj.size(1)
        if _fused_lstm_ok(self, self.encoder, obs_traj):
            return ops.lstm_encoder(obs_traj, self.spatial_embedding, self.encoder)
        if _fused_lstm_train_ok(self, self.encoder, obs_traj):
            return ops.lstm_encoder_train(obs_traj, self.spatial_embedding, self.encoder)
        emb = self.spatial_embedding(obs_traj.reshape(-1, 2)).view(-1, batch, self.embedding_dim)
        _, state = self.encoder(emb, self.init_hidden(batch, emb))
        return state[0]


class Decoder(nn.Module):
    """LSTM step loop; optional per-step pooling with its own PoolHiddenNet (sgan/models.py:95-178)."""

    def __init__(self, seq_len, embedding_dim=64, h_dim=128, mlp_dim=1024, num_layers=1, pool_every_timestep=True,
                 dropout=0.0, bottleneck_dim=1024, activation='relu', batch_norm=True, pooling_type='pool_net',
                 neighborhood_size=2.0, grid_size=8):
        super().__init__()
        self.seq_len = seq_len
        self.mlp_dim = mlp_dim
        self.h_dim = h_dim
        self.embedding_dim = embedding_dim
        self.pool_every_timestep = pool_every_timestep
        self.spatial_embedding = nn.Linear(2, embedding_dim)
        self.decoder = nn.LSTM(embedding_dim, h_dim, num_layers, dropout=dropout)
        self.hidden2pos = nn.Linear(h_dim, 2)
        if pool_every_timestep:
            if pooling_type == 'pool_net':
                self.pool_net = PoolHiddenNet(embedding_dim=embedding_dim, h_dim=h_dim, mlp_dim=mlp_dim,
                                              bottleneck_dim=bottleneck_dim, activation=activation,
                                              batch_norm=batch_norm, dropout=dropout)
            self.mlp = make_mlp([h_dim + bottleneck_dim, mlp_dim, h_dim], activation=activation,
                                batch_norm=batch_norm, dropout=dropout)

    def forward(self, last_pos, last_pos_rel, state_tuple, seq_start_end):
        batch = last_pos.size(0)
        if _fused_lstm_ok(self, self.decoder, last_pos_rel, state_tuple[0], state_tuple[1]):
            return self._forward_fused(last_pos, last_pos_rel, state_tuple, seq_start_end)
        if not self.pool_every_timestep and _fused_lstm_train_ok(self, self.decoder, last_pos_rel, *state_tuple):
            pred, hf = ops.lstm_decoder_train(state_tuple[0], state_tuple[1], last_pos_rel, self.seq_len,
                                              self.spatial_embedding, self.decoder, self.hidden2pos)
            return pred, hf.unsqueeze(0)
        steps = []
        dec_in = self.spatial_embedding(last_pos_rel).view(1, batch, self.embedding_dim)
        for _ in range(self.seq_len):
            output, state_tuple = self.decoder(dec_in, state_tuple)
            rel_pos = self.hidden2pos(output.view(-1, self.h_dim))
            curr_pos = rel_pos + last_pos
            if self.pool_every_timestep:
                dec_h = state_tuple[0]
                pool_h = self.pool_net(dec_h, seq_start_end, curr_pos)
                dec_h = self.mlp(torch.cat([dec_h.view(-1, self.h_dim), pool_h], dim=1))
                state_tuple = (dec_h.unsqueeze(0), state_tuple[1])
            dec_in = self.spatial_embedding(rel_pos).view(1, batch, self.embedding_dim)
            steps.append(rel_pos.view(batch, -1))
            last_pos = curr_pos
        return torch.stack(steps, dim=0), state_tuple[0]


def _decoder_forward_fused(self, last_pos, last_pos_rel, state_tuple, seq_start_end):
    """Inference path of Decoder.forward: the whole step loop in one kernel (no per-step pooling), or one fused
    cell + hidden2pos kernel per step around the pooling op (pool_every_timestep)."""
    h, c = state_tuple
    if not self.pool_every_timestep:
        pred, hf = ops.lstm_decoder(h, c, last_pos_rel, self.seq_len, self.spatial_embedding, self.decoder,
                                    self.hidden2pos, want_state='h')
        return pred, hf.unsqueeze(0)
    h, c = h.reshape(-1, self.h_dim), c.reshape(-1, self.h_dim)
    rel_in, steps = last_pos_rel, []
    for _ in range(self.seq_len):
        rel, h, c = ops.lstm_decoder(h, c, rel_in, 1, self.spatial_embedding, self.decoder, self.hidden2pos,
                                     want_state=True)
        rel = rel[0]
        curr_pos = rel + last_pos
        pool_h = self.pool_net(h, seq_start_end, curr_pos)
        fused = ops.mlp2(self.mlp, h, pool_h)                       # cat + mlp (models.py:165-166) in one launch
        h = fused if fused is not None else self.mlp(torch.cat([h, pool_h], dim=1))
        steps.append(rel)
        rel_in, last_pos = rel, curr_pos
    return torch.stack(steps, dim=0), h.unsqueeze(0)


Decoder._forward_fused = _decoder_forward_fused


class TrajectoryGenerator(nn.Module):
    def __init__(self, obs_len, pred_len, embedding_dim=64, encoder_h_dim=64, decoder_h_dim=128, mlp_dim=1024,
                 num_layers=1, noise_dim=(0,), noise_type='gaussian', noise_mix_type='ped', pooling_type=None,
                 pool_every_timestep=True, dropout=0.0, bottleneck_dim=1024, activation='relu', batch_norm=True,
                 neighborhood_size=2.0, grid_size=8, n_units=(32, 16, 32), n_heads=4, dropout1=0, alpha=0.2,
                 context_type='gat'):
        super().__init__()
        if pooling_type and pooling_type.lower() == 'none':
            pooling_type = None
        if context_type not in ('gat', 'gcn', 'mlp'):
            raise ValueError('context_type must be gat | gcn | mlp')
        self.obs_len = obs_len
        self.pred_len = pred_len
        self.mlp_dim = mlp_dim
        self.encoder_h_dim = encoder_h_dim
        self.decoder_h_dim = decoder_h_dim
        self.embedding_dim = embedding_dim
        self.noise_dim = noise_dim
        self.num_layers = num_layers
        self.noise_type = noise_type
        self.noise_mix_type = noise_mix_type
        self.pooling_type = pooling_type
        self.noise_first_dim = 0
        self.pool_every_timestep = pool_every_timestep
        self.bottleneck_dim = 1024
        self.context_type = context_type

        self.encoder = Encoder(embedding_dim=embedding_dim, h_dim=encoder_h_dim, mlp_dim=mlp_dim,
                               num_layers=num_layers, dropout=dropout)
        self.decoder = Decoder(pred_len, embedding_dim=embedding_dim, h_dim=decoder_h_dim, mlp_dim=mlp_dim,
                               num_layers=num_layers, pool_every_timestep=pool_every_timestep, dropout=dropout,
                               bottleneck_dim=bottleneck_dim, activation=activation, batch_norm=batch_norm,
                               pooling_type=pooling_type, grid_size=grid_size, neighborhood_size=neighborhood_size)
        if pooling_type == 'pool_net':
            self.pool_net = PoolHiddenNet(embedding_dim=embedding_dim, h_dim=encoder_h_dim, mlp_dim=mlp_dim,
                                          bottleneck_dim=bottleneck_dim, activation=activation, batch_norm=batch_norm)
        if self.noise_dim is None or self.noise_dim[0] == 0:
            self.noise_dim = None
        else:
            self.noise_first_dim = noise_dim[0]

        input_dim = encoder_h_dim + bottleneck_dim if pooling_type else encoder_h_dim
        ctx_out = decoder_h_dim - self.noise_first_dim
        if context_type == 'gat':
            # live reference: both modules exist, only gatencoder is called (gcn_module is a passenger)
            self.gatencoder = GATEncoder(n_units=n_units, n_heads=n_heads, dropout=dropout1, alpha=alpha)
            self.gcn_module = GCNModule(input_dim=input_dim, hidden_dim=72, out_dim=16, gcn_layers=2, final_dim=ctx_out)
        elif context_type == 'gcn':
            self.gcn_module = GCNModule(input_dim=input_dim, hidden_dim=72, out_dim=16, gcn_layers=2, final_dim=ctx_out)
            self.mlp_decoder_context = make_mlp([input_dim, mlp_dim, ctx_out], activation=activation,
                                                batch_norm=batch_norm, dropout=dropout)
        else:
            self.mlp_decoder_context = make_mlp([input_dim, mlp_dim, ctx_out], activation=activation,
                                                batch_norm=batch_norm, dropout=dropout)

    def add_noise(self, _input, seq_start_end, user_noise=None):
        if not self.noise_dim:
            return _input
        sched = get_schedule(seq_start_end, _input.device)
        if self.noise_mix_type == 'global':
            noise_shape = (sched.n_scenes,) + tuple(self.noise_dim)
        else:
            noise_shape = (_input.size(0),) + tuple(self.noise_dim)
        z = user_noise if user_noise is not None else get_noise(noise_shape, self.noise_type, _input.device)
        z = z.to(_input.device)
        if self.noise_mix_type == 'global':
            z = z.reshape(sched.n_scenes, -1).index_select(0, ped_scene_index(sched))
        return torch.cat([_input, z], dim=1)

    def mlp_decoder_needed(self):
        return bool(self.noise_dim or self.pooling_type or self.encoder_h_dim != self.decoder_h_dim)

    def context(self, obs_traj, obs_traj_rel, seq_start_end, obs_traj_g):
        """Everything of forward() that does not depend on the noise sample (encoder, pooling, graph context)."""
        final_encoder_h = self.encoder(ready(obs_traj_rel))      # (ready: staged host batches, utils.stage_host_batch)
        end_pos = ready_last(obs_traj)[-1, :, :]
        if self.pooling_type:
            pool_h = self.pool_net(final_encoder_h, seq_start_end, end_pos)
            if self.context_type == 'mlp' and self.mlp_decoder_needed():
                # SGAN-P wiring (models.py:886,898): cat + the two Linear+ReLU layers as one launch at inference
                fused = ops.mlp2(self.mlp_decoder_context, final_encoder_h.view(-1, self.encoder_h_dim), pool_h)
                if fused is not None:
                    return fused
            ctx_in = torch.cat([final_encoder_h.view(-1, self.encoder_h_dim), pool_h], dim=1)
        else:
            ctx_in = final_encoder_h.view(-1, self.encoder_h_dim)
        if not self.mlp_decoder_needed():
            return ctx_in
        if self.context_type == 'gat':
            return self.gatencoder(ctx_in, seq_start_end, end_pos, ready_last(obs_traj_g)[-1, :, :])
        if self.context_type == 'gcn':
            return self.gcn_module(ctx_in, seq_start_end, end_pos, ready_last(obs_traj_g)[-1, :, :])
        fused = ops.mlp2(self.mlp_decoder_context, ctx_in)
        return fused if fused is not None else self.mlp_decoder_context(ctx_in)

    def decode(self, ctx, obs_traj, obs_traj_rel, seq_start_end, user_noise=None):
        ready_last(obs_traj), ready(obs_traj_rel)          # (only obs_traj[-1] is read below)
        batch = obs_traj_rel.size(1)
        dec = self.decoder
        if (self.noise_dim and self.noise_mix_type == 'global' and not dec.pool_every_timestep and len(self.noise_dim) == 1
                and _fused_lstm_ok(dec, dec.decoder, ctx, obs_traj_rel)):
            # inference fast path: add_noise (models.py:837-846) is folded into the fused decoder kernel
            sched = get_schedule(seq_start_end, ctx.device)
            z = user_noise if user_noise is not None else get_noise((sched.n_scenes,) + tuple(self.noise_dim),
                                                                   self.noise_type, ctx.device)
            z = z.to(ctx.device).reshape(sched.n_scenes, -1)
            return ops.lstm_decoder(ctx, None, obs_traj_rel[-1], dec.seq_len, dec.spatial_embedding, dec.decoder,
                                    dec.hidden2pos, z=z, ped_scene=sched.ped_scene32())
        decoder_h = self.add_noise(ctx, seq_start_end, user_noise=user_noise).unsqueeze(0)
        decoder_c = decoder_h.new_zeros(self.num_layers, batch, self.decoder_h_dim)
        pred_rel, _ = self.decoder(obs_traj[-1], obs_traj_rel[-1], (decoder_h, decoder_c), seq_start_end)
        return pred_rel

    def forward(self, obs_traj, obs_traj_rel, seq_start_end, obs_traj_g, user_noise=None):
        ctx = self.context(obs_traj, obs_traj_rel, seq_start_end, obs_traj_g)
        return self.decode(ctx, obs_traj, obs_traj_rel, seq_start_end, user_noise=user_noise)


def ped_scene_index(sched):
    """int64 [batch]: scene index of every pedestrian (cached on the schedule)."""
    idx = getattr(sched, '_ped_scene', None)
    if idx is None:
        counts = (sched.scene_start[1:] - sched.scene_start[:-1]).long()
        idx = torch.repeat_interleave(torch.arange(sched.n_scenes, device=counts.device), counts,
                                      output_size=sched.batch)
        sched._ped_scene = idx
    return idx


class TrajectoryDiscriminator(nn.Module):
    def __init__(self, obs_len, pred_len, embedding_dim=64, h_dim=64, mlp_dim=1024, num_layers=1, activation='relu',
                 batch_norm=True, dropout=0.0, d_type='local'):
        super().__init__()
        self.obs_len = obs_len
        self.pred_len = pred_len
        self.seq_len = obs_len + pred_len
        self.h_dim = h_dim
        self.d_type = d_type
        self.encoder = Encoder(embedding_dim=embedding_dim, h_dim=h_dim, mlp_dim=mlp_dim, num_layers=num_layers,
                               dropout=dropout)
        if d_type == 'global':
            self.pool_net = PoolHiddenNet(embedding_dim=embedding_dim, h_dim=h_dim, mlp_dim=mlp_dim,
                                          bottleneck_dim=h_dim, activation=activation, batch_norm=batch_norm)
        self.real_classifier = make_mlp([h_dim, mlp_dim, 1], activation=activation, batch_norm=batch_norm,
                                        dropout=dropout)

    def forward(self, traj, traj_rel, seq_start_end=None):
        final_h = self.encoder(traj_rel)
        if self.d_type == 'local':
            classifier_input = final_h.squeeze()
        else:
            classifier_input = self.pool_net(final_h.squeeze(), seq_start_end, traj[0])
        fused = ops.mlp2(self.real_classifier, classifier_input) if classifier_input.dim() == 2 else None
        return fused if fused is not None else self.real_classifier(classifier_input)
