import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vidtok_b200 import _native as N
torch.zeros(1).cuda()
buf = ctypes.create_string_buffer(512)
for smem in (50_000, 100_000, 150_000, 200_000, 215_000, 230_000):
    n = N.lib().vt_debug_cluster_query(smem, buf, 512)
    print(smem, n, buf.value.decode())
