"""Timing of the skinny (M = 20) decoder GEMMs: weight-streaming mma.sync kernel (gemm_skinny.cu) vs the tcgen05 tile kernels.
Graph replay; `cold` cycles through 8 weight copies (> L2) so the weights come from HBM, `hot` re-reads one copy (L2-resident)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dasa_b200 import lib, ops

L = lib.load()
M = 20
for (N, K, epi) in [(4096, 2240, ops.EPI_NONE), (4096, 1024, ops.EPI_NONE), (2176, 1024, ops.EPI_NONE), (2048, 1024, ops.EPI_NONE),
                    (1024, 3072, ops.EPI_TANH), (1024, 4096, ops.EPI_NONE), (2240, 4096, ops.EPI_NONE), (1024, 2176, ops.EPI_NONE),
                    (64, 128, ops.EPI_BIAS_TANH)]:
    ncopy = 8 if N * K * 4 * 8 > 200e6 else max(8, int(200e6 / (N * K * 4)) + 1)
    Ws = [torch.randn(N, K, device="cuda") * 0.05 for _ in range(ncopy)]
    A = torch.randn(M, K, device="cuda"); b = torch.randn(N, device="cuda"); C = torch.empty(M, N, device="cuda")
    out = []
    for mode in (0, 2):
        L.dasa_debug_gemm_skinny(mode)
        row = []
        for cold in (False, True):
            kw = dict(epilogue=epi, precision=ops.PREC_TF32)
            if epi == ops.EPI_BIAS_TANH:
                kw["bias"] = b
            for W in Ws[:2]:
                ops.gemm(A, K, 1, W, K, 1, C, N, M, N, K, **kw)
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            reps = 4 * ncopy
            with torch.cuda.graph(g):
                for i in range(reps):
                    ops.gemm(A, K, 1, Ws[i % ncopy] if cold else Ws[0], K, 1, C, N, M, N, K, **kw)
            g.replay(); torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
            e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
            row.append(e0.elapsed_time(e1) * 1e3 / reps)
        out.append(row)
    gb = N * K * 4 / 1e3
    print("M=20 N=%5d K=%5d  tile kernels: hot %6.2f us cold %6.2f us | skinny: hot %6.2f us (%5.0f GB/s) cold %6.2f us (%5.0f GB/s)" % (
        N, K, out[0][0], out[0][1], out[1][0], gb / out[1][0], out[1][1], gb / out[1][1]), flush=True)
L.dasa_debug_gemm_skinny(1)
