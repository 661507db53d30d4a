"""Timing of the MN-major operand layouts of the CTA-pair TF32 GEMM against the K-major kernel fed transposed copies (and the
cost of those copies) on the backward shapes of the rollout. usage: python scripts/pair_mn_gemm.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dasa_b200 import lib, ops

L = lib.load()


def timeit(fn, reps=10):
    fn(); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps):
            fn()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / reps


def main():
    # dW[N, K] += dY[M, N]^T X[M, K]   (a_kmajor = b_kmajor = 0: reduction over the M rows)
    for (M, N, K, what) in [(35700, 4096, 768, "bi-LSTM dW_ih"), (35700, 4096, 1024, "bi-LSTM dW_hh"), (32900, 2048, 2048, "AdaIN dW"),
                            (700, 4096, 2240, "decoder dW_ih"), (700, 4096, 1024, "decoder dW_hh")]:
        dY = torch.randn(M, N, device="cuda"); X = torch.randn(M, K, device="cuda"); dW = torch.zeros(N, K, device="cuda")
        if not L.dasa_gemm_layout_on_tensor_cores(0, 0, N, K, M):
            print("%-16s M=%d N=%d K=%d: not on the pair kernel" % (what, M, N, K)); continue
        t_mn = timeit(lambda: ops.gemm(dY, N, 0, X, K, 0, dW, K, N, K, M, beta=1.0, precision=ops.PREC_TF32))
        dYt, Xt = ops.transpose(dY), ops.transpose(X)
        t_k = timeit(lambda: ops.gemm(dYt, dYt.stride(0), 1, Xt, Xt.stride(0), 1, dW, K, N, K, M, beta=1.0, precision=ops.PREC_TF32))
        t_tr = timeit(lambda: (ops.transpose(dY), ops.transpose(X)))
        fl = 2.0 * M * N * K
        print("%-16s rows=%6d dW %dx%d : MN-major %7.1f us %6.1f TF/s | K-major %7.1f us %6.1f TF/s + transposes %7.1f us" % (
            what, M, N, K, t_mn, fl / t_mn / 1e6, t_k, fl / t_k / 1e6, t_tr), flush=True)
    # dX[M, K] = dY[M, N] W[N, K]   (B MN-major)
    for (M, N, K, what) in [(35700, 4096, 768, "bi-LSTM dX"), (20300, 3072, 768, "FFN dX (finetune)"), (25200, 768, 2176, "visn_fc dX")]:
        dY = torch.randn(M, N, device="cuda"); W = torch.randn(N, K, device="cuda") * 0.05; dX = torch.empty(M, K, device="cuda")
        if not L.dasa_gemm_layout_on_tensor_cores(1, 0, M, K, N):
            print("%-16s: not on the pair kernel" % what); continue
        t_mn = timeit(lambda: ops.gemm(dY, N, 1, W, K, 0, dX, K, M, K, N, precision=ops.PREC_TF32))
        Wt = ops.transpose(W)
        t_k = timeit(lambda: ops.gemm(dY, N, 1, Wt, Wt.stride(0), 1, dX, K, M, K, N, precision=ops.PREC_TF32))
        fl = 2.0 * M * N * K
        print("%-16s M=%6d N=%5d K=%5d: B MN-major %7.1f us %6.1f TF/s | K-major %7.1f us %6.1f TF/s" % (
            what, M, N, K, t_mn, fl / t_mn / 1e6, t_k, fl / t_k / 1e6), flush=True)


if __name__ == "__main__":
    main()
