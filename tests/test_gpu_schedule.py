"""The scene schedule derived on the device (csrc/sgx_schedule.cu: scan over scenes + binary searches per pedestrian and
per 128-pair tile) against the host pass (sgx_schedule_build): integer work, bit exact, on ragged layouts from one lone
pedestrian to 2^16 scenes, dense crowds, and the K-tiled schedules of the folded forwards."""
import numpy as np
import pytest
import torch

from conftest import sse_from_sizes

pytestmark = pytest.mark.gpu

DEV = 'cuda:0'


def _both(sse):
    from group_gan_gcn_gat_b200 import _lib
    from group_gan_gcn_gat_b200.schedule import SceneSchedule
    prev = _lib.option('sched_device')
    try:
        _lib.set_option('sched_device', 0)
        host = SceneSchedule(sse, DEV)
        _lib.set_option('sched_device', 1)
        dev = SceneSchedule(sse, DEV)
    finally:
        _lib.set_option('sched_device', prev)
    return host, dev


def _layouts():
    rng = np.random.RandomState(0)
    yield 'one pedestrian', [1]
    yield 'one scene', [57]
    yield 'pair of scenes', [2, 3]
    yield 'tile boundary', [8, 8, 11, 1, 16]           # 64 + 64 = exactly one 128-pair tile, then 121 + 1 + 256
    yield 'dense crowd', [1024]
    yield 'dense crowds', [1024, 3, 700, 1, 1, 512]
    yield 'lone pedestrians', [1] * 5000
    yield 'zara-like 1k', list(rng.choice([1, 2, 3, 4, 5, 6, 8, 11, 17, 26], size=1000))
    yield 'eth-like 65k', list(rng.choice([1, 2, 3, 4, 5, 7, 9, 13, 21, 40], size=1 << 16,
                                          p=[.2, .2, .15, .12, .1, .08, .06, .05, .03, .01]))
    yield 'scan block boundaries', list(rng.randint(1, 6, size=4 * 1024 + 1))
    yield 'mixed 300k scenes', list(rng.randint(1, 4, size=300000))


@pytest.mark.parametrize('name,sizes', list(_layouts()), ids=[n for n, _ in _layouts()])
def test_device_built_schedule_is_bit_equal_to_the_host_pass(name, sizes):
    sse = sse_from_sizes(sizes)
    host, dev = _both(sse)
    assert (host.batch, host.n_pairs, host.n_tiles, host.max_n) == (dev.batch, dev.n_pairs, dev.n_tiles, dev.max_n)
    for field in ('pair_off', 'scene_start', 'ped_start', 'ped_end', 'tile_first'):
        assert torch.equal(getattr(host, field), getattr(dev, field)), field
    assert torch.equal(host.ped_scene32(), dev.ped_scene32())
    for cap in (32, 64):
        (ch, nh), (cd, nd) = host.chunks(cap), dev.chunks(cap)
        assert nh == nd and torch.equal(ch[:nh + 1].cpu(), cd[:nd + 1].cpu()), 'chunks(%d)' % cap


def test_device_built_tiled_schedule_matches_host():
    from group_gan_gcn_gat_b200 import _lib
    from group_gan_gcn_gat_b200.schedule import tiled_schedule
    sse = sse_from_sizes([3, 1, 9, 2, 40, 5])
    prev = _lib.option('sched_device')
    try:
        _lib.set_option('sched_device', 0)
        host = tiled_schedule(sse.clone(), 20, DEV)
        _lib.set_option('sched_device', 1)
        dev = tiled_schedule(sse.clone(), 20, DEV)
    finally:
        _lib.set_option('sched_device', prev)
    for field in ('pair_off', 'scene_start', 'ped_start', 'ped_end', 'tile_first'):
        assert torch.equal(getattr(host, field), getattr(dev, field)), field


def test_device_build_rejects_inconsistent_totals():
    from group_gan_gcn_gat_b200 import _lib
    L = _lib.lib()
    buf = torch.zeros(1 << 16, dtype=torch.uint8, device=DEV)
    p = buf.data_ptr()
    rc = L.sgx_schedule_build_device(p, 4, 10, 30, 7, p, p, p, p, p, p, p, 1 << 16, None)      # n_tiles != ceil(30 / 128)
    assert rc != 0 and b'sgx_schedule_stats' in L.sgx_last_error()
