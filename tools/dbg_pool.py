import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tools.time_pool import time_case
G = (16, 32, 8)
for mode in ("ts",):
    os.environ['SGX_POOL_TC_MODE'] = mode
    for dbg in (0, 1, 3, 15):
        os.environ['SGX_POOL_TC_DBG'] = str(dbg)
        print(mode, 'dbg', dbg, end='  ')
        time_case([1024] * 8, G, 'bf16', reps=5)
