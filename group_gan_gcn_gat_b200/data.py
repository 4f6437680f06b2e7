"""Input pipeline of the hot path: the group-labelled trajectory dataset and its collate.

Mirror of sgan/data/trajectories_GCN.py (TrajectoryDataset :77-204, seq_collate :15-42, read_file :45-56,
poly_fit :59-74) and sgan/data/loader.py (data_loader :9-29) -- same constructor arguments, attributes, item
layout and batch tuple, so scripts/train.py:152 and scripts/evaluate_model.py:115 work unchanged:

    dset, loader = data_loader(args, path)
    (obs_traj, pred_traj_gt, obs_traj_rel, pred_traj_gt_rel, obs_vel, pred_vel,
     obs_traj_g, pred_traj_g, non_linear_ped, loss_mask, seq_start_end) = batch

What is different is how the tensors are produced.  The reference scans every window of every file with nested
Python loops over frames and pedestrians (7.4 s for zara1/train).  Here one file is one sort plus a handful of array
operations: rows are ordered by (ped, frame, file position), and a pedestrian is complete in the window starting at
frame index w exactly when L consecutive rows of that order span frame indices w .. w+L-1 with no other row of that
pedestrian inside the window (that is the reference's `pad_end - pad_front == seq_len and len == seq_len` test,
:139-143).  All complete (window, ped) pairs are gathered at once, the quadratic fits of poly_fit are one batched
least-squares call, and windows with <= min_ped complete pedestrians are dropped (:161).

`DeviceLoader` is the B200-side replacement of DataLoader + seq_collate for the training / evaluation loops: it
draws the same batches in the same order as `DataLoader(dset, shuffle=..., collate_fn=seq_collate)` under the same
torch seed (it uses torch's own samplers), builds each batch with one index gather per tensor instead of per-sample
slicing, stages it in pinned memory and copies it to the device on a side stream one batch ahead of the consumer.
"""
import hashlib
import logging
import math
import os

import numpy as np
import torch
from torch.utils.data import DataLoader, Dataset

logger = logging.getLogger(__name__)

BATCH_FIELDS = ('obs_traj', 'pred_traj', 'obs_traj_rel', 'pred_traj_rel', 'obs_vel', 'pred_vel', 'obs_traj_g',
                'pred_traj_g', 'non_linear_ped', 'loss_mask', 'seq_start_end')


def seq_collate(data):
    """list of dataset items -> the 11-tuple of the reference (trajectories_GCN.py:15-42): sequence tensors
    [seq_len, batch, C], non_linear_ped [batch], loss_mask [batch, seq_len], seq_start_end LongTensor [S, 2]."""
    cols = list(zip(*data))
    counts = [len(seq) for seq in cols[0]]
    ends = np.cumsum(counts).tolist()
    starts = [0] + ends[:-1]
    out = [torch.cat(c, dim=0).permute(2, 0, 1) for c in cols[:8]]
    out.append(torch.cat(cols[8]))
    out.append(torch.cat(cols[9], dim=0))
    out.append(torch.LongTensor([[s, e] for s, e in zip(starts, ends)]))
    return tuple(out)


def read_file(_path, delim='\t'):
    """<frame_id> <ped_id> <x> <y> <group label> per line -> float64 [rows, 5].  Like the reference (:45-56) the
    separator is always a tab: its `delim` argument is normalised and then ignored."""
    del delim
    rows = []
    with open(_path, 'r') as f:
        for line in f:
            rows.append(line.strip().split('\t'))
    return np.asarray(rows, dtype=np.float64)


def poly_fit(traj, traj_len, threshold):
    """1.0 when the last traj_len points of a [2+, T] trajectory leave a quadratic-fit residual >= threshold
    (trajectories_GCN.py:59-74)."""
    flags = _poly_fit_batch(np.asarray(traj)[None, :2, -traj_len:], threshold)
    return float(flags[0])


def _poly_fit_batch(xy, threshold):
    """xy [C, 2, n] -> float64 [C] of 0/1 flags; one np.polyfit call with 2C right-hand sides."""
    C, _, n = xy.shape
    if C == 0:
        return np.zeros(0)
    t = np.linspace(0, n - 1, n)
    res = np.polyfit(t, xy.reshape(2 * C, n).T, 2, full=True)[1]
    if res.size == 0:                      # rank-deficient / exactly determined fit: numpy reports no residual
        return np.zeros(C)
    res = res.reshape(C, 2)
    return (res[:, 0] + res[:, 1] >= threshold).astype(np.float64)


def scan_file(data, obs_len, pred_len, skip=1, threshold=0.002, min_ped=1):
    """All kept windows of one file.  -> dict(seq [P,2,L], rel [P,2,L], g [P,1,L], non_linear [P], counts [W])
    in the reference's order: windows by ascending start frame, pedestrians by ascending id."""
    L = obs_len + pred_len
    empty = dict(seq=np.zeros((0, 2, L)), rel=np.zeros((0, 2, L)), g=np.zeros((0, 1, L)), non_linear=np.zeros(0),
                 counts=np.zeros(0, dtype=np.int64))
    n = data.shape[0]
    if n < L:
        return empty
    frames, fi = np.unique(data[:, 0], return_inverse=True)
    fi = fi.reshape(-1)
    num_sequences = int(math.ceil((len(frames) - L + 1) / skip))
    last_start = num_sequences * skip            # range(0, num_sequences * skip + 1, skip), :118
    if last_start < 0:
        return empty
    order = np.lexsort((np.arange(n), fi, data[:, 1]))
    ped, f = data[order, 1], fi[order]
    i = np.arange(n - L + 1)
    ok = (ped[i + L - 1] == ped[i]) & (f[i + L - 1] - f[i] == L - 1)
    ok[1:] &= ~((ped[i[1:] - 1] == ped[i[1:]]) & (f[i[1:] - 1] == f[i[1:]]))      # an earlier row inside the window
    j = i[:-1] + L
    ok[:-1] &= ~((ped[j] == ped[i[:-1]]) & (f[j] == f[j - 1]))                     # a later row inside the window
    ok &= (f[i] % skip == 0) & (f[i] <= last_start)
    cand = i[ok]
    if cand.size == 0:
        return empty
    assert data.shape[1] - 2 == 3, 'dataset has no labeling'                      # trajectories_GCN.py:153
    w, pid = f[cand], ped[cand]
    by = np.lexsort((pid, w))
    cand, w = cand[by], w[by]
    rows = order[cand[:, None] + np.arange(L)[None, :]]
    vals = np.around(data[rows][:, :, 2:], decimals=4).transpose(0, 2, 1)         # [C, 3, L]
    non_linear = _poly_fit_batch(vals[:, :2, L - pred_len:], threshold)
    starts, counts = np.unique(w, return_counts=True)
    keep_w = counts > min_ped
    keep = np.repeat(keep_w, counts)
    seq = np.ascontiguousarray(vals[keep, :2, :])
    rel = np.zeros_like(seq)
    rel[:, :, 1:] = seq[:, :, 1:] - seq[:, :, :-1]
    return dict(seq=seq, rel=rel, g=np.ascontiguousarray(vals[keep, 2:, :]), non_linear=non_linear[keep],
                counts=counts[keep_w].astype(np.int64))


class TrajectoryDataset(Dataset):
    """Dataloder for the Trajectory datasets with group labels (trajectories_GCN.py:77-204): same arguments, same
    attributes (obs_traj [P,2,obs_len], pred_traj, *_rel, *_g [P,1,*], loss_mask [P,L], non_linear_ped [P],
    seq_start_end list of (start, end), num_seq), same item layout.

    cache_dir: optional directory for the tensorised result, keyed by the file list (name, size, mtime) and the
    arguments -- the "cache the tensorised dataset" half of SURVEY 8f row f4.
    """

    def __init__(self, data_dir, obs_len=8, pred_len=12, skip=1, threshold=0.002, min_ped=1, delim='\t',
                 cache_dir=None):
        super().__init__()
        self.data_dir = data_dir
        self.obs_len = obs_len
        self.pred_len = pred_len
        self.skip = skip
        self.seq_len = obs_len + pred_len
        self.delim = delim
        all_files = [os.path.join(data_dir, p) for p in os.listdir(data_dir)]       # listdir order, like :101-102
        arrays = self._load_cache(cache_dir, all_files, threshold, min_ped)
        if arrays is None:
            parts = [scan_file(read_file(p, delim), obs_len, pred_len, skip, threshold, min_ped) for p in all_files]
            if not any(len(p['counts']) for p in parts):
                raise ValueError('need at least one array to concatenate')          # what np.concatenate([]) says, :170
            arrays = {k: np.concatenate([p[k] for p in parts], axis=0) for k in ('seq', 'rel', 'g', 'non_linear', 'counts')}
            self._save_cache(cache_dir, arrays)
        seq, rel, g, counts = arrays['seq'], arrays['rel'], arrays['g'], arrays['counts']
        self.num_seq = int(len(counts))
        o = obs_len
        as_f32 = lambda a: torch.from_numpy(np.ascontiguousarray(a)).type(torch.float)  # noqa: E731
        self.obs_traj, self.pred_traj = as_f32(seq[:, :, :o]), as_f32(seq[:, :, o:])
        self.obs_traj_rel, self.pred_traj_rel = as_f32(rel[:, :, :o]), as_f32(rel[:, :, o:])
        self.obs_traj_g, self.pred_traj_g = as_f32(g[:, :, :o]), as_f32(g[:, :, o:])
        # every kept pedestrian is present in all seq_len frames (pad_front = 0, pad_end = seq_len, :139-141)
        self.loss_mask = torch.ones(seq.shape[0], self.seq_len, dtype=torch.float)
        self.non_linear_ped = as_f32(arrays['non_linear'])
        ends = np.cumsum(counts).tolist()
        self.seq_start_end = [(s, e) for s, e in zip([0] + ends[:-1], ends)]

    # -- cache ------------------------------------------------------------------------------------------------
    def _cache_key(self, files, threshold, min_ped):
        h = hashlib.sha256()
        for p in files:
            st = os.stat(p)
            h.update(('%s|%d|%d;' % (os.path.basename(p), st.st_size, st.st_mtime_ns)).encode())
        h.update(repr((self.obs_len, self.pred_len, self.skip, threshold, min_ped)).encode())
        return h.hexdigest()[:24]

    def _load_cache(self, cache_dir, files, threshold, min_ped):
        self._cache_path = None
        if not cache_dir:
            return None
        self._cache_path = os.path.join(cache_dir, 'trajectories_%s.npz' % self._cache_key(files, threshold, min_ped))
        if not os.path.isfile(self._cache_path):
            return None
        with np.load(self._cache_path) as z:
            return {k: z[k] for k in z.files}

    def _save_cache(self, cache_dir, arrays):
        if self._cache_path:
            os.makedirs(cache_dir, exist_ok=True)
            tmp = self._cache_path + '.tmp%d.npz' % os.getpid()
            np.savez(tmp, **arrays)
            os.replace(tmp, self._cache_path)

    # -- Dataset protocol -------------------------------------------------------------------------------------
    def __len__(self):
        return self.num_seq

    def __getitem__(self, index):
        start, end = self.seq_start_end[index]
        return [
            self.obs_traj[start:end, :], self.pred_traj[start:end, :],
            self.obs_traj_rel[start:end, :], self.pred_traj_rel[start:end, :],
            self.obs_traj_rel[start:end, :] * 2.5, self.pred_traj_rel[start:end, :] * 2.5,     # velocity = rel / 0.4 s
            self.obs_traj_g[start:end, :], self.pred_traj_g[start:end, :],
            self.non_linear_ped[start:end], self.loss_mask[start:end, :],
        ]

    def collate_indices(self, indices):
        """seq_collate([self[i] for i in indices]) with one gather per tensor."""
        se = torch.as_tensor([self.seq_start_end[int(i)] for i in indices], dtype=torch.int64).reshape(-1, 2)
        counts = se[:, 1] - se[:, 0]
        ends = torch.cumsum(counts, 0)
        starts = ends - counts
        total = int(ends[-1]) if len(indices) else 0
        rows = torch.arange(total) - torch.repeat_interleave(starts - se[:, 0], counts)
        take = lambda t: t.index_select(0, rows)                                            # noqa: E731
        seqs = [take(t).permute(2, 0, 1) for t in (self.obs_traj, self.pred_traj, self.obs_traj_rel, self.pred_traj_rel)]
        out = seqs + [seqs[2] * 2.5, seqs[3] * 2.5,
                      take(self.obs_traj_g).permute(2, 0, 1), take(self.pred_traj_g).permute(2, 0, 1),
                      take(self.non_linear_ped), take(self.loss_mask), torch.stack([starts, ends], 1)]
        return tuple(out)


def data_loader(args, path):
    """(dataset, DataLoader) exactly as sgan/data/loader.py:9-29 builds them (shuffle=True, seq_collate)."""
    dset = TrajectoryDataset(path, obs_len=args.obs_len, pred_len=args.pred_len, skip=args.skip, delim=args.delim,
                             cache_dir=getattr(args, 'dataset_cache_dir', None))
    loader = DataLoader(dset, batch_size=args.batch_size, shuffle=True, num_workers=args.loader_num_workers,
                        collate_fn=seq_collate)
    return dset, loader


class DeviceLoader:
    """Iterates the same batches, in the same order, as DataLoader(dset, batch_size, shuffle, collate_fn=seq_collate)
    under the same torch seed, but hands them out already on `device`: batch k+1 is gathered, staged in pinned host
    memory and copied on a side stream while the consumer works on batch k.  `device=None` keeps batches on the host
    (gather only).  len() and drop_last follow DataLoader."""

    def __init__(self, dset, batch_size=64, shuffle=True, device=None, drop_last=False, generator=None):
        self.dset, self.batch_size, self.shuffle, self.drop_last = dset, batch_size, shuffle, drop_last
        self.device = torch.device(device) if device is not None else None
        self.generator = generator
        if self.device is not None and self.device.type != 'cuda':
            raise ValueError('DeviceLoader stages batches for a CUDA device; got %s' % (self.device,))

    def __len__(self):
        n = len(self.dset)
        return n // self.batch_size if self.drop_last else (n + self.batch_size - 1) // self.batch_size

    def _batches(self):
        # a real DataLoader over the bare indices: it consumes the torch RNG exactly like the reference's loader does
        # (base seed first, then the sampler's permutation), whatever the torch version
        return DataLoader(range(len(self.dset)), batch_size=self.batch_size, shuffle=self.shuffle,
                          drop_last=self.drop_last, generator=self.generator, collate_fn=list)

    def __iter__(self):
        if self.device is None:
            for idx in self._batches():
                yield self.dset.collate_indices(idx)
            return
        side = torch.cuda.Stream(self.device)
        main = torch.cuda.current_stream(self.device)

        def stage(idx):
            host = [t.contiguous().pin_memory() for t in self.dset.collate_indices(idx)]
            with torch.cuda.stream(side):
                dev = tuple(t.to(self.device, non_blocking=True) for t in host)
                done = torch.cuda.Event()
                done.record(side)
            return dev, done, host                       # `host` stays referenced until the copy has been consumed

        pending = None
        for idx in self._batches():
            nxt = stage(idx)
            if pending is not None:
                dev, done, _host = pending
                main.wait_event(done)
                for t in dev:
                    t.record_stream(main)
                yield dev
            pending = nxt
        if pending is not None:
            dev, done, _host = pending
            main.wait_event(done)
            for t in dev:
                t.record_stream(main)
            yield dev
