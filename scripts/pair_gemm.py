"""Correctness + timing of the persistent CTA-pair TF32 GEMM (gemm_tc2.cu) against the single-CTA kernel and torch.
usage: python scripts/pair_gemm.py [check|time]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dasa_b200 import lib, ops

L = lib.load()
torch.backends.cuda.matmul.allow_tf32 = False


def run(M, N, K, epi, mode, bias=None, reps=0):
    A = torch.randn(M, K, device="cuda"); W = torch.randn(N, K, device="cuda") * 0.05
    b = torch.randn(N, device="cuda") if bias is None else bias
    C = torch.empty(M, N, device="cuda")
    L.dasa_debug_gemm_pair(mode)
    kw = dict(epilogue=epi, precision=ops.PREC_TF32)
    if epi in (ops.EPI_BIAS, ops.EPI_BIAS_GELU, ops.EPI_BIAS_TANH, ops.EPI_BIAS_RELU):
        kw["bias"] = b
    ops.gemm(A, K, 1, W, K, 1, C, N, M, N, K, **kw)
    torch.cuda.synchronize()
    us = None
    if reps:
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for _ in range(reps):
                ops.gemm(A, K, 1, W, K, 1, C, N, M, N, K, **kw)
        g.replay(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / reps
    return A, W, b, C, us


def check():
    ok = True
    for (M, N, K) in [(256, 256, 32), (777, 515, 96), (256, 256, 64), (512, 512, 256), (300, 200, 100), (20300, 768, 768), (2500, 2304, 768), (1000, 132, 40)]:
        for epi in (ops.EPI_NONE, ops.EPI_BIAS, ops.EPI_BIAS_GELU):
            A, W, b, C, _ = run(M, N, K, epi, 2)
            ref = A.double() @ W.double().t()
            if epi != ops.EPI_NONE:
                ref = ref + b.double()
            if epi == ops.EPI_BIAS_GELU:
                ref = torch.nn.functional.gelu(ref)
            err = float((C.double() - ref).abs().max() / ref.abs().max())
            print("M=%d N=%d K=%d epi=%d rel err %.2e" % (M, N, K, epi, err), flush=True)
            ok &= err < 3e-3
    print("CHECK", "OK" if ok else "FAILED")


def time_():
    for (M, N, K) in [(20300, 3072, 768), (20300, 768, 3072), (20300, 2304, 768), (20300, 768, 768), (25200, 768, 2176),
                      (25200, 768, 768), (25200, 3072, 768), (35000, 2048, 2048), (20300, 8192, 768), (8192, 8192, 8192),
                      (2320, 3072, 768), (2320, 768, 3072), (2320, 2304, 768), (4640, 768, 768), (4096, 4096, 4096), (14848, 3072, 768)]:
        res = []
        for mode in (0, 2):
            _, _, _, _, us = run(M, N, K, ops.EPI_BIAS_GELU if N == 3072 else ops.EPI_BIAS, mode, reps=10)
            res.append(us)
        fl = 2.0 * M * N * K
        print("M=%6d N=%5d K=%5d  single-CTA %8.1f us %6.1f TF/s | pair256 %8.1f us %6.1f TF/s" % (
            M, N, K, res[0], fl / res[0] / 1e6, res[1], fl / res[1] / 1e6), flush=True)
    L.dasa_debug_gemm_pair(1)


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "check"
    check() if what == "check" else time_()
