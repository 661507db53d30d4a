"""GPU helper: the fp16-operand CTA-pair GEMM (dasa_gemm_f16) on the frozen stack's shapes, graph-free, L2 flushed between launches."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dasa_b200 import ops
ops.set_precision("tf32")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def timeit(fn, n=8):
    ts = []
    for i in range(n):
        flush.zero_()
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


shapes = [(20300, 3072, 768, ops.EPI_BIAS_GELU, True), (20300, 3072, 768, ops.EPI_BIAS, True), (20300, 768, 3072, ops.EPI_BIAS, True),
          (20300, 768, 3072, ops.EPI_BIAS, False), (20300, 2304, 768, ops.EPI_BIAS, True), (20300, 768, 768, ops.EPI_BIAS, True),
          (20300, 768, 768, ops.EPI_BIAS, False), (25200, 1536, 768, ops.EPI_BIAS, True), (25200, 768, 768, ops.EPI_BIAS, True),
          (25200, 3072, 768, ops.EPI_BIAS_GELU, True), (25200, 768, 3072, ops.EPI_BIAS, True), (25200, 2304, 768, ops.EPI_BIAS, True)]
for M, N, K, epi, half in shapes:
    x = (torch.randn(M, K, device="cuda") * 0.5).half()
    w = (torch.randn(N, K, device="cuda") * 0.05).half()
    b = torch.randn(N, device="cuda")
    ms = timeit(lambda: ops.linear_f16(x, w, b, epi, half))
    byt = 2 * (M * K + N * K) + (2 if half else 4) * M * N
    print("f16 gemm %6d x %5d x %5d epi=%d out=%s: %7.1f us  %7.1f TFLOP/s  %6.0f GB/s" % (
        M, N, K, epi, "f16" if half else "f32", ms * 1e3, 2.0 * M * N * K / ms / 1e9, byt / ms / 1e6))
