"""Small helpers of the reference that the wiring needs (sgan/utils.py:83-96)."""
import torch


def relative_to_abs(rel_traj, start_pos):
    """[T,batch,2] displacements + [batch,2] start -> absolute positions [T,batch,2]."""
    return torch.cumsum(rel_traj, dim=0) + start_pos.unsqueeze(0)


# ---- host batches: staged copies with one event per tensor -------------------------------------------------------
# scripts/evaluate_model.py:75 moves the whole minibatch to the GPU (`[tensor.cuda() for tensor in batch]`) before the
# first forward.  The generator does not need all of it at once: the encoder reads obs_traj_rel only, pooling and the
# metrics the LAST step of obs_traj, the graph context the last step of obs_traj_g, the metrics pred_traj_gt -- so the
# copies go to a copy stream in that order, each followed by an event, and the compute stream waits for a tensor where
# it is first read (`ready` / `ready_last`): the first sample's encoder runs while most of the batch is still on the
# PCIe link, and the earlier steps of obs_traj / obs_traj_g (which nothing on the device reads) arrive last.
_copy_streams = {}


def _copy_stream(device):
    side = _copy_streams.get(device)
    if side is None:
        side = _copy_streams[device] = torch.cuda.Stream(device)
    return side


def stage_host_batch(device, obs_traj, obs_traj_rel, obs_traj_g, pred_traj_gt):
    """Host (ideally pinned) tensors of one minibatch -> device tensors whose copies are in flight on a copy stream, in
    the order the forward reads them.  Every returned tensor carries the events of its own copy: `ready(t)` makes the
    current stream wait for all of it, `ready_last(t)` for its last time step only.  The host tensors must stay unchanged
    until the results of the step have been read."""
    device = torch.device(device)
    main = torch.cuda.current_stream(device)
    side = _copy_stream(device)
    srcs = (obs_traj, obs_traj_rel, obs_traj_g, pred_traj_gt)
    outs = [torch.empty(t.shape, dtype=t.dtype, device=device) for t in srcs]       # blocks of the compute stream's pool
    obs_d, rel_d, grp_d, gt_d = outs
    side.wait_stream(main)             # the blocks may have been freed by work that is still queued on the compute stream

    def copy(dst, src, attr, owner):
        if dst.numel():
            dst.copy_(src, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record(side)
        setattr(owner, attr, ev)

    with torch.cuda.stream(side):
        copy(rel_d, obs_traj_rel, '_sgx_ready', rel_d)
        copy(obs_d[-1], obs_traj[-1], '_sgx_ready_last', obs_d)
        copy(grp_d[-1], obs_traj_g[-1], '_sgx_ready_last', grp_d)
        copy(gt_d, pred_traj_gt, '_sgx_ready', gt_d)
        copy(obs_d[:-1], obs_traj[:-1], '_sgx_ready', obs_d)
        copy(grp_d[:-1], obs_traj_g[:-1], '_sgx_ready', grp_d)
    for t in outs:
        t.record_stream(side)          # a tensor nobody waited for must not be recycled under its own copy
    return obs_d, rel_d, grp_d, gt_d


def ready(t):
    """Make the current stream wait for the whole staged copy of `t` (stage_host_batch); a no-op for every other tensor
    and after the first call.  Returns t."""
    for attr in ('_sgx_ready_last', '_sgx_ready'):
        ev = getattr(t, attr, None)
        if ev is not None:
            torch.cuda.current_stream(t.device).wait_event(ev)
            setattr(t, attr, None)
    return t


def ready_last(t):
    """As `ready`, for readers of t[-1] only (the last observed step: end_pos, the group labels of the graph context)."""
    if not hasattr(t, '_sgx_ready_last'):
        return ready(t)                       # not staged step-split (or not staged at all: a no-op)
    ev = t._sgx_ready_last
    if ev is not None:
        torch.cuda.current_stream(t.device).wait_event(ev)
        t._sgx_ready_last = None
    return t
