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
        self.host_sse = sse
        self.device = torch.device(device)
        self._chunks = {}
        # All index arrays live in ONE buffer: filled on the host (pinned staging when the target is a GPU), uploaded
        # with ONE async copy.  Seven separate pageable copies cost 1.4 ms per minibatch at 65 k scenes, as much as
        # the H2D of the trajectories themselves.
        B, T = self.batch, max(self.n_tiles, 1)
        fused_chunks = self.max_n <= 32
        fields = [('pair_off', np.int64, B + 1), ('scene_start', np.int32, S + 1), ('ped_start', np.int32, B),
                  ('ped_end', np.int32, B), ('tile_first', np.int32, T), ('ped_scene', np.int32, B),
                  ('chunk_scene', np.int32, S + 1 if fused_chunks else 0)]
        offs, total = {}, 0
        for name, dt, n in fields:
            offs[name] = total
            total += (n * np.dtype(dt).itemsize + 15) // 16 * 16
        on_gpu = self.device.type == 'cuda'
        if on_gpu and _lib.option('sched_device'):
            # Derived on the GPU (csrc/sgx_schedule.cu): the host has validated seq_start_end and taken its totals (one pass
            # over the SCENES, above); it uploads 16 bytes per scene and four small launches write the per-pedestrian and
            # per-tile arrays -- the host pass over the pedestrians and the upload of its arrays (0.75 ms at 65 k scenes,
            # in front of the first pooling launch of a minibatch) are gone.  The chunk list of the graph kernels is a
            # sequential greedy packing: built on the host when a graph kernel first asks for it (`chunks`).
            # seq_start_end goes into a pinned staging buffer and the first launch reads it from there (mapped host memory):
            # no copy-engine transfer, which would queue behind the H2D prefetch of the next minibatch.
            ws_off = (total + 255) // 256 * 256
            ws_bytes = int(L.sgx_schedule_device_ws_bytes(S))
            host_buf, done = _staging(S * 16, self.device)
            np.frombuffer(host_buf.numpy(), dtype=np.int64, count=2 * S)[:] = sse.reshape(-1)
            dev_buf = torch.empty(ws_off + ws_bytes, dtype=torch.uint8, device=self.device)
            stream = torch.cuda.current_stream(self.device)
            base = dev_buf.data_ptr()
            with torch.cuda.device(self.device):
                _lib.check(L.sgx_schedule_build_device(host_buf.data_ptr(), S, self.batch, self.n_pairs, self.n_tiles,
                                                       base + offs['scene_start'], base + offs['ped_start'],
                                                       base + offs['ped_end'], base + offs['pair_off'],
                                                       base + offs['tile_first'], base + offs['ped_scene'],
                                                       base + ws_off, ws_bytes, stream.cuda_stream), 'sgx_schedule_build_device')
            done.record(stream)              # the staging buffer is free again once the first launch has read it
            self._finish(dev_buf, offs, B, S, T, None)
            return
        host_buf, done = _staging(total, self.device) if on_gpu else (torch.empty(max(total, 16), dtype=torch.uint8), None)
        raw = host_buf.numpy()
        view = {name: np.frombuffer(raw, dtype=dt, count=n, offset=offs[name]) for name, dt, n in fields}
        n = np.zeros(1, np.int64)
        _lib.check(L.sgx_schedule_build(sse.ctypes.data, S, view['scene_start'].ctypes.data, view['ped_start'].ctypes.data,
                                        view['ped_end'].ctypes.data, view['pair_off'].ctypes.data,
                                        view['tile_first'].ctypes.data, view['ped_scene'].ctypes.data, 32,
                                        view['chunk_scene'].ctypes.data if fused_chunks else None, n.ctypes.data),
                   'seq_start_end')
        n_chunks = int(n[0])
        if on_gpu:
            dev_buf = torch.empty(max(total, 16), dtype=torch.uint8, device=self.device)
            dev_buf.copy_(host_buf[:max(total, 16)], non_blocking=True)
            done.record(torch.cuda.current_stream(self.device))
        else:
            dev_buf = host_buf
        self._finish(dev_buf, offs, B, S, T, n_chunks if fused_chunks else -1)

    def _finish(self, dev_buf, offs, B, S, T, n_chunks):
        """typed views into the one device buffer; n_chunks: count of the chunk list in the buffer, -1: a scene is larger
        than 32 (no list), None: not built yet (device-built schedule: `chunks` builds it on first use)"""
        self._buf = dev_buf

        def dev(name, dt, n):
            width = np.dtype(dt).itemsize
            return dev_buf[offs[name]:offs[name] + n * width].view(torch.int64 if dt is np.int64 else torch.int32)
        self.pair_off = dev('pair_off', np.int64, B + 1)
        self.scene_start = dev('scene_start', np.int32, S + 1)
        self.ped_start = dev('ped_start', np.int32, B)
        self.ped_end = dev('ped_end', np.int32, B)
        self.tile_first = dev('tile_first', np.int32, T)
        self._ped_scene32 = dev('ped_scene', np.int32, B)
        if n_chunks is not None:
            self._chunks[32] = (dev('chunk_scene', np.int32, n_chunks + 1), n_chunks) if n_chunks >= 0 else \
                (self.scene_start[:0], 0)

    def ped_scene32(self):
        """int32 [batch] scene index of every pedestrian (device), for the noise fold-in of the fused decoder."""
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
                hit = (_upload_i32(buf[:int(n[0]) + 1], self.device), int(n[0]))
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


_staging_bufs = {}   # device -> (pinned uint8 tensor, event of the last upload that read it)


def _staging(nbytes, device, slot='sched'):
    """Grow-only pinned staging buffer per device (and use); waits for the previous upload out of it before it is refilled."""
    key = (device.type, device.index, slot)
    hit = _staging_bufs.get(key)
    if hit is not None:
        hit[1].synchronize()
    if hit is None or hit[0].numel() < nbytes:
        buf = torch.empty(max(int(nbytes * 1.5), 1 << 16), dtype=torch.uint8).pin_memory()
        hit = (buf, torch.cuda.Event())
        _staging_bufs[key] = hit
    return hit


def _upload_i32(arr, device):
    """small int32 host array -> device tensor through its own pinned staging buffer, asynchronously and without a
    copy-engine transfer (a pageable `.to(device)` synchronises the host with the stream in the middle of a step)"""
    if device.type != 'cuda':
        return torch.from_numpy(np.ascontiguousarray(arr).copy())
    n16 = (int(arr.size) * 4 + 15) // 16 * 16
    host_buf, done = _staging(max(n16, 16), device, slot='i32')
    np.frombuffer(host_buf.numpy(), dtype=np.int32, count=arr.size)[:] = arr
    out = torch.empty(n16 // 4, dtype=torch.int32, device=device)
    stream = torch.cuda.current_stream(device)
    with torch.cuda.device(device):            # fetched by a kernel from the mapped staging buffer (see SceneSchedule)
        _lib.check(_lib.lib().sgx_fetch_pinned(out.data_ptr(), host_buf.data_ptr(), n16, stream.cuda_stream), 'sgx_fetch_pinned')
    done.record(stream)
    return out[:arr.size]


_cache = {}   # id(tensor) -> (weakref to tensor, version key, schedule); Tensor.__eq__ rules out WeakKeyDictionary


def _evict_dead():
    for k in [k for k, v in _cache.items() if v[0]() is None]:
        del _cache[k]


def tiled_schedule(seq_start_end, k, device):
    """Schedule of k copies of the batch laid side by side (copy c covers rows c * batch + [start, end) of every scene): what
    the K-folded forwards and the stacked fake + real discriminator batch run on.  Built from the HOST copy of the base
    schedule and cached on it -- `seq_start_end.repeat(k, 1) + offsets` on the device made every folded step read the new
    tensor back (a device synchronisation in the middle of a launch-bound training step)."""
    base = get_schedule(seq_start_end, device)
    tiled = base.__dict__.setdefault('_tiled', {})
    hit = tiled.get(k)
    if hit is None:
        big = (base.host_sse[None] + (np.arange(k, dtype=np.int64) * base.batch)[:, None, None]).reshape(-1, 2)
        hit = tiled[k] = SceneSchedule(big, device)
    return hit


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
        # drop the schedules of batches that no longer exist BEFORE building the new one: their device buffers go back to
        # the caching allocator and are reused right away.  (Evicting only every 64 misses let the reserved memory of an
        # evaluation loop grow by one schedule per minibatch, i.e. a cudaMalloc -- 10 .. 150 ms of host stall -- every
        # few steps.)
        _evict_dead()
        sched = SceneSchedule(seq_start_end, device)
        _cache[id(seq_start_end)] = (weakref.ref(seq_start_end), key, sched)
        return sched
    return SceneSchedule(seq_start_end, device)
