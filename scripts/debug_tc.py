"""GPU debug helper: tcgen05 TF32 GEMM vs float64, with a per-tile error map when something is off."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dasa_b200 import ops

def run(M, N, K, epi=ops.EPI_NONE, beta=0.0, lda_pad=0):
    g = torch.Generator().manual_seed(M * 7 + N * 3 + K)
    A = torch.randn(M, K + lda_pad, generator=g)[:, :K]
    W = torch.randn(N, K, generator=g) / K ** 0.5
    b = torch.randn(N, generator=g)
    C0 = torch.randn(M, N, generator=g)
    ref = A.double() @ W.double().t()
    if epi == ops.EPI_BIAS: ref = ref + b.double()
    ref = ref + beta * C0.double()
    Ad, Wd, C = A.cuda(), W.cuda(), C0.cuda().clone()
    if lda_pad:
        Ad = torch.randn(M, K + lda_pad, generator=g).cuda(); Ad[:, :K] = A.cuda(); Ad = Ad[:, :K]
    ops.gemm(Ad, Ad.stride(0), 1, Wd, K, 1, C, N, M, N, K, beta=beta, epilogue=epi, bias=b.cuda(), precision=ops.PREC_TF32)
    torch.cuda.synchronize()
    err = (C.double().cpu() - ref).abs()
    rel = float(err.max() / ref.abs().max())
    print("M=%d N=%d K=%d epi=%d beta=%g: max rel err %.3e  (mean abs err %.3e, ref rms %.3e)" % (M, N, K, epi, beta, rel, float(err.mean()), float(ref.pow(2).mean().sqrt())), flush=True)
    if rel > 5e-3:
        bm = err.reshape(-1)[: (M // 32) * 32 * N] if False else None
        tiles = []
        for i in range(0, M, 32):
            row = []
            for j in range(0, N, 32):
                row.append("%.1e" % float(err[i:i + 32, j:j + 32].max()))
            tiles.append(" ".join(row[:12]))
        print("\n".join(tiles[:12]))
    return rel

if __name__ == "__main__":
    print(torch.cuda.get_device_name(0))
    shapes = [(128, 128, 32), (128, 128, 64), (128, 128, 256), (128, 64, 256), (256, 256, 512), (1600, 768, 768), (720, 768, 2176),
              (1600, 2304, 768), (1600, 3072, 768), (1600, 768, 3072), (20, 4096, 3264), (20, 2176, 1024), (1000, 2048, 2048),
              (100, 200, 300), (37, 53, 64), (129, 65, 100)]
    for s in shapes:
        try:
            run(*s)
        except Exception as e:
            print("FAILED", s, repr(e)[:300], flush=True)
            break
    run(300, 500, 700, ops.EPI_BIAS, 0.0)
    run(300, 500, 700, ops.EPI_NONE, 1.0)
    run(300, 512, 700, ops.EPI_NONE, 0.0, lda_pad=4)
    # timing
    for (M, N, K) in [(1600, 2304, 768), (1600, 3072, 768), (1600, 768, 3072), (720, 768, 2176), (1000, 2048, 2048), (20, 4096, 3264), (8192, 8192, 8192)]:
        A = torch.randn(M, K, device="cuda"); W = torch.randn(N, K, device="cuda"); C = torch.empty(M, N, device="cuda")
        for prec in (ops.PREC_TF32, ops.PREC_FP32):
            if prec == ops.PREC_FP32 and M * N * K > 1e11: continue
            for _ in range(3): ops.gemm(A, K, 1, W, K, 1, C, N, M, N, K, precision=prec)
            e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
            e0.record()
            for _ in range(10): ops.gemm(A, K, 1, W, K, 1, C, N, M, N, K, precision=prec)
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 10
            print("time M=%d N=%d K=%d prec=%d: %.3f ms  %.1f TFLOP/s  (weights %.1f GB/s)" % (M, N, K, prec, ms, 2 * M * N * K / ms / 1e9, N * K * 4 / ms / 1e6), flush=True)
        torch.backends.cuda.matmul.allow_tf32 = True
        for _ in range(3): torch.matmul(A, W.t())
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        for _ in range(10): torch.matmul(A, W.t())
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        print("   cuBLAS tf32 reference: %.3f ms  %.1f TFLOP/s" % (ms, 2 * M * N * K / ms / 1e9), flush=True)
