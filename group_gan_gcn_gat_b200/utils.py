"""Small helpers of the reference that the wiring needs (sgan/utils.py:83-96)."""
import torch


def relative_to_abs(rel_traj, start_pos):
    """[T,batch,2] displacements + [batch,2] start -> absolute positions [T,batch,2]."""
    return torch.cumsum(rel_traj, dim=0) + start_pos.unsqueeze(0)
