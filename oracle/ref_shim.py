"""Import the UNMODIFIED reference from /root/reference on CPU (build container only).

TEST INFRASTRUCTURE: used by oracle/make_golden.py (fixture generation) and by the
optional ``tests/test_oracle_vs_reference.py`` which is skipped when /root/reference
is absent (it is absent on the GPU box).  Nothing in the product imports this.

Shims (all outside the reference tree):
  1. ``.cuda()`` -> identity           (hard-coded .cuda() at sgan/models.py:26,58,267,...)
  2. stub ``attrdict`` module          (scripts/evaluate_model.py:11)
  3. ``torch.load`` -> map_location='cpu', weights_only=False
"""
import os
import sys
import types

import torch

REF_ROOT = os.environ.get('SGAN_REFERENCE_ROOT', '/root/reference')


def available():
    return os.path.isfile(os.path.join(REF_ROOT, 'sgan', 'models.py'))


_loaded = {}


def load():
    """Returns the reference ``sgan.models`` module (import side effects applied once)."""
    if 'models' in _loaded:
        return _loaded['models']
    if not available():
        raise RuntimeError('reference tree not present at %s' % REF_ROOT)
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    torch.Tensor.cuda = lambda self, *a, **k: self
    torch.nn.Module.cuda = lambda self, *a, **k: self
    stub = types.ModuleType('attrdict')

    class AttrDict(dict):
        def __getattr__(self, k):
            try:
                return self[k]
            except KeyError:
                raise AttributeError(k)

    stub.AttrDict = AttrDict
    sys.modules.setdefault('attrdict', stub)
    import sgan.models as ref_models  # noqa: E402
    _loaded['models'] = ref_models
    return ref_models


def load_checkpoint(rel_path):
    return torch.load(os.path.join(REF_ROOT, rel_path), map_location='cpu', weights_only=False)


def load_dataset(name, split, obs_len=8, pred_len=12, batch_size=64):
    """The reference's own loader (sgan/data/loader.py:9-29) with shuffle disabled for reproducibility."""
    load()
    from sgan.data.trajectories_GCN import TrajectoryDataset, seq_collate
    from torch.utils.data import DataLoader
    path = os.path.join(REF_ROOT, 'datasets_group', name, split)
    dset = TrajectoryDataset(path, obs_len=obs_len, pred_len=pred_len, skip=1, delim='tab')
    return dset, DataLoader(dset, batch_size=batch_size, shuffle=False, num_workers=0, collate_fn=seq_collate)
