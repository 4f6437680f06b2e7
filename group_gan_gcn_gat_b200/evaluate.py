"""Best-of-K evaluation of one minibatch -- the loop body of scripts/evaluate_model.py:72-99 without per-scene
python loops (SURVEY 8f rows f2/f3 stay behind the same API: every sample is a complete generator forward).

    ade_sum, fde_sum = evaluate_batch(generator, obs_traj, obs_traj_rel, seq_start_end, obs_traj_g, pred_traj_gt, K)

ade_sum / (total_traj * pred_len) and fde_sum / total_traj are the numbers the reference prints.
"""
import torch

from .losses import displacement_error, final_displacement_error
from .models import ped_scene_index
from .schedule import get_schedule
from .utils import relative_to_abs


def best_of_k_sum(per_sample, sched):
    """evaluate_helper (scripts/evaluate_model.py:58-69): per scene sum over peds, min over the K samples, summed."""
    seg = ped_scene_index(sched)
    per_scene = per_sample.new_zeros(sched.n_scenes, per_sample.shape[1]).index_add_(0, seg, per_sample)
    return per_scene.min(dim=1).values.sum()


@torch.no_grad()
def evaluate_batch(generator, obs_traj, obs_traj_rel, seq_start_end, obs_traj_g, pred_traj_gt, num_samples=20,
                   noise=None, hoist_context=False):
    """hoist_context=True computes the noise-independent part of the forward (encoder, pooling, graph context:
    everything before sgan/models.py:909) once instead of num_samples times -- bit-identical results when the
    decoder does not pool per step.  Off by default: the reference recomputes it per sample."""
    sched = get_schedule(seq_start_end, obs_traj.device)
    if noise is None and generator.noise_dim and generator.noise_mix_type == 'global':
        # one draw for all K samples on the device generator.  The reference draws each sample on the CPU generator and
        # copies it (sgan/models.py:23-29), which costs more host time than the whole forward here; pass `noise=` to
        # reproduce a CPU-seeded stream.
        fn = torch.randn if generator.noise_type == 'gaussian' else (lambda *a, **k: torch.rand(*a, **k) * 2 - 1)
        noise = fn(num_samples, sched.n_scenes, *generator.noise_dim, device=obs_traj.device)
    dev = obs_traj.device
    if obs_traj.is_cuda and num_samples <= 32:
        # fused path: one metrics kernel per sample, one best-of-K kernel (sgx_displacement_errors / sgx_best_of_k)
        from . import _lib
        from .ops import _f32, _ptr, _stream
        L = _lib.lib()
        batch, T = obs_traj.shape[1], pred_traj_gt.shape[0]
        ade = torch.empty(batch, num_samples, dtype=torch.float32, device=dev)
        fde = torch.empty_like(ade)
        out2 = torch.empty(2, dtype=torch.float32, device=dev)
        gt, start = _f32(pred_traj_gt, 'pred_traj_gt'), _f32(obs_traj[-1], 'obs_traj')
        ctx = generator.context(obs_traj, obs_traj_rel, seq_start_end, obs_traj_g) if hoist_context else None
        with torch.cuda.device(dev):
            for k in range(num_samples):
                nz = None if noise is None else noise[k]
                if ctx is not None:
                    rel = generator.decode(ctx, obs_traj, obs_traj_rel, seq_start_end, user_noise=nz).contiguous()
                else:
                    rel = generator(obs_traj, obs_traj_rel, seq_start_end, obs_traj_g, user_noise=nz).contiguous()
                _lib.check(L.sgx_displacement_errors(_ptr(rel), _ptr(start), _ptr(gt), T, batch, _ptr(ade), _ptr(fde),
                                                     num_samples, k, _stream(rel)), 'sgx_displacement_errors')
            _lib.check(L.sgx_best_of_k(_ptr(ade), _ptr(fde), _ptr(sched.scene_start), sched.n_scenes, num_samples,
                                       _ptr(out2), _stream(ade)), 'sgx_best_of_k')
        return out2[0], out2[1]
    ade, fde = [], []
    for k in range(num_samples):
        rel = generator(obs_traj, obs_traj_rel, seq_start_end, obs_traj_g,
                        user_noise=None if noise is None else noise[k])
        pred = relative_to_abs(rel, obs_traj[-1])
        ade.append(displacement_error(pred, pred_traj_gt, mode='raw'))
        fde.append(final_displacement_error(pred[-1], pred_traj_gt[-1], mode='raw'))
    return best_of_k_sum(torch.stack(ade, dim=1), sched), best_of_k_sum(torch.stack(fde, dim=1), sched)
