"""GATEncoder / GCNModule operator timings at the bench size (2^16 zara1-shaped scenes): the forward at inference (tcgen05
kernel, group structure in-kernel) and forward + backward under autograd.  The same numbers bench.py prints as
hot_path_ops, without the rest of the bench."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

dev = torch.device('cuda:0')
torch.cuda.set_device(dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
rows = bench.hot_path_op_numbers(dev, 1 << 16, 1.0, flush)
for r in rows:
    print('%-60s %8.4f ms' % (r['op'][:60], r['ms']))
