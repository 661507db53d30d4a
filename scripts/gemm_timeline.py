import sys, os, ctypes
os.environ["DASA_TC_DEBUG"] = "2"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dasa_b200 import ops, lib
L = lib.load()
def t(M, N, K):
    A = torch.randn(M, K, device="cuda"); W = torch.randn(N, K, device="cuda"); b = torch.randn(N, device="cuda"); C = torch.empty(M, N, device="cuda")
    for _ in range(5): ops.gemm(A, K, 1, W, K, 1, C, N, M, N, K, epilogue=ops.EPI_BIAS, bias=b, precision=ops.PREC_TF32)
    torch.cuda.synchronize()
    buf = (ctypes.c_ulonglong * 16)()
    L.dasa_debug_tc_timestamps.argtypes = [ctypes.c_void_p]
    L.dasa_debug_tc_timestamps(buf)
    ts = [int(x) for x in buf[:9]]
    names = ["start", "prologue done", "1st TMA issued", "all TMA issued", "1st full", "MMA all issued+commit", "tmem_full seen", "epilogue done", "dealloc"]
    print("M=%d N=%d K=%d (last CTA), ns since start:" % (M, N, K))
    for n, x in zip(names, ts): print("   %-24s %8d" % (n, x - ts[0]))
t(896, 2304, 32); t(896, 2304, 768); t(128, 128, 768)
