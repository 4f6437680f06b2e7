"""Scene schedule: seq_start_end -> flat per-ped index arrays, built once per minibatch.

The reference walks ``for (start, end) in seq_start_end`` with two ``.item()`` device syncs per scene
in every module call (sgan/models.py:507-510, 256-262, 639-644).  Here the ragged layout is read once
(one D2H copy if the tensor lives on the GPU), turned into flat int arrays by the C ABI
(``sgx_schedule_fill``) and cached on the tensor object, so the K best-of-K samples, the pred_len decoder
steps and G/D all reuse it without touching the host again.
"""
import weakref

import numpy as np
import torch

from . import _lib


class SceneSchedule:
    """Device-resident description of the ragged scenes of one minibatch."""

    def __init__(self, seq_start_end, device):
        sse = seq_start_end
        if torch.is_tensor(sse):
            sse = sse.detach().to('cpu', torch.int64).contiguous().numpy()
        sse = np.ascontiguousarray(np.asarray(sse, dtype=np.int64).reshape(-1, 2))
        L = _lib.lib()
        S = int(sse.shape[0])
        stats = np.zeros(8, np.int64)
        _lib.check(L.sgx_schedule_stats(sse.ctypes.data, S, stats.ctypes.data), 'seq_start_end')
        self.n_scenes = S
        self.batch = int(stats[0])
        self.max_n = int(stats[1])
        self.n_pairs = int(stats[2])
        self.n_tiles = int(stats[3])
        scene_start = np.empty(S + 1, np.int32)
        ped_start = np.empty(self.batch, np.int32)
        ped_end = np.empty(self.batch, np.int32)
        pair_off = np.empty(self.batch + 1, np.int64)
        tile_first = np.empty(max(self.n_tiles, 1), np.int32)
        _lib.check(L.sgx_schedule_fill(sse.ctypes.data, S, scene_start.ctypes.data, ped_start.ctypes.data,
                                       ped_end.ctypes.data, pair_off.ctypes.data, tile_first.ctypes.data),
                   'seq_start_end')
        self.host_sse = sse
        self.device = torch.device(device)
        put = lambda a: torch.from_numpy(a).to(self.device, non_blocking=False)
        self.scene_start = put(scene_start)
        self.ped_start = put(ped_start)
        self.ped_end = put(ped_end)
        self.pair_off = put(pair_off)
        self.tile_first = put(tile_first)
        self._chunks = {}
        self._ped_scene32 = None

    def ped_scene32(self):
        """int32 [batch] scene index of every pedestrian (device), for the noise fold-in of the fused decoder."""
        if self._ped_scene32 is None:
            sizes = (self.host_sse[:, 1] - self.host_sse[:, 0]).astype(np.int64)
            idx = np.repeat(np.arange(self.n_scenes, dtype=np.int32), sizes)
            self._ped_scene32 = torch.from_numpy(idx).to(self.device)
        return self._ped_scene32

    def chunks(self, cap=32):
        """(chunk_scene int32 device tensor, n_chunks) packing whole scenes into chunks of <= cap peds, or (empty, 0)
        when a scene is larger than cap (the caller then takes the multi-pass path).  Cached."""
        hit = self._chunks.get(cap)
        if hit is None:
            if self.max_n > cap:
                hit = (torch.empty(0, dtype=torch.int32, device=self.device), 0)
            else:
                buf = np.empty(self.n_scenes + 1, np.int32)
                n = np.zeros(1, np.int64)
                _lib.check(_lib.lib().sgx_schedule_chunks(self.host_sse.ctypes.data, self.n_scenes, cap,
                                                          buf.ctypes.data, n.ctypes.data), 'chunks')
                hit = (torch.from_numpy(buf[:int(n[0]) + 1].copy()).to(self.device), int(n[0]))
            self._chunks[cap] = hit
        return hit

    def partition(self, world):
        """LPT split of scenes over ranks by N^2 cost -> (rank_of_scene int32 [S], cost per rank)."""
        L = _lib.lib()
        rank = np.empty(self.n_scenes, np.int32)
        cost = np.zeros(world, np.int64)
        _lib.check(L.sgx_schedule_partition(self.host_sse.ctypes.data, self.n_scenes, world, rank.ctypes.data,
                                            cost.ctypes.data), 'partition')
        return rank, cost


_cache = {}   # id(tensor) -> (weakref to tensor, version key, schedule); Tensor.__eq__ rules out WeakKeyDictionary


def _evict_dead():
    for k in [k for k, v in _cache.items() if v[0]() is None]:
        del _cache[k]


def get_schedule(seq_start_end, device):
    """Cached per seq_start_end tensor object (and its in-place version counter)."""
    device = torch.device(device)
    if isinstance(seq_start_end, SceneSchedule):
        return seq_start_end
    if torch.is_tensor(seq_start_end):
        key = (seq_start_end._version, device.type, device.index)
        hit = _cache.get(id(seq_start_end))
        if hit is not None and hit[0]() is seq_start_end and hit[1] == key:
            return hit[2]
        sched = SceneSchedule(seq_start_end, device)
        if len(_cache) > 64:
            _evict_dead()
        _cache[id(seq_start_end)] = (weakref.ref(seq_start_end), key, sched)
        return sched
    return SceneSchedule(seq_start_end, device)
