import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dasa_b200 import lib
L = lib.load()
ctr = torch.zeros(32, dtype=torch.int32, device="cuda")
out = torch.zeros(1, dtype=torch.int64, device="cuda")
junk = torch.zeros(256 * 160, device="cuda")
st = torch.cuda.current_stream().cuda_stream
for variant in (0, 1, 2, 3, 4):
    for it in range(3):
        rc = L.dasa_debug_barrier_bench(variant, 1000, ctypes.c_void_p(ctr.data_ptr()), ctypes.c_void_p(out.data_ptr()), ctypes.c_void_p(junk.data_ptr()), ctypes.c_void_p(st))
        torch.cuda.synchronize()
        print("variant %d: rc=%d %.3f us per barrier (at 1965 MHz)" % (variant, rc, int(out[0]) / 1000 / 1965.0))
