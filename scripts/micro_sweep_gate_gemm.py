"""BASELINE.json configs[4], GEMM-fused form of K1: out = sigmoid(d W^T + b) * f (* keep mask), the gate saved on the side —
dasa_gemm with DASA_EPI_GATE on strided [B, 36, C+128] feature buffers, against the TF32 tensor roofline (measured bf16 / 2).
Also the a2 backward kernel (channel_modulate_bwd) against the HBM roofline. L2 flushed between launches, median of 7."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dasa_b200 import ops
ops.set_precision("tf32")
pk = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))
peak_tf, peak_gb = pk["bf16_tflops_sustained"] / 2.0, pk["hbm_gbs"]
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def timeit(fn, reps=7):
    ts = []
    for i in range(reps + 2):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        if i >= 2:
            ts.append(e0.elapsed_time(e1) * 1e-3)
    ts.sort()
    return ts[len(ts) // 2]


V, A = 36, 128
print("# K1 GEMM-fused sigmoid gate (tcgen05 TF32 + fused epilogue), C, B, rows, us, TFLOP/s, %% of TF32 peak (%.0f)" % peak_tf)
for C in (2048, 4096):
    F = C + A
    W = torch.randn(C, C, device="cuda") / C ** 0.5
    b = torch.randn(C, device="cuda") * 0.02
    for B in (1, 16, 64, 256, 700, 1024, 2048):
        R = B * V
        f = torch.rand(B, V, F, device="cuda"); d = torch.rand(B, V, F, device="cuda"); o = torch.empty(B, V, F, device="cuda")
        s = torch.empty(R, C, device="cuda")
        keep = (torch.rand(R, C, device="cuda") >= 0.4).to(torch.uint8)
        f2, d2, o2 = f.view(R, F), d.view(R, F), o.view(R, F)
        fn = lambda: ops.gemm(d2, F, 1, W, C, 1, o2, F, R, C, C, epilogue=ops.EPI_GATE, bias=b, gate_src=f2, ld_gate=F, gate_out=s,
                              ld_gate_out=C, drop_mask=keep, drop_scale=1 / 0.6)
        t = timeit(fn)
        tf = 2.0 * R * C * C / t / 1e12
        print("K1 gate GEMM  C=%4d B=%5d rows=%6d %9.1f us %7.1f TFLOP/s %5.1f%%" % (C, B, R, t * 1e6, tf, 100 * tf / peak_tf), flush=True)
        del f, d, o, s, keep
print("# a2 backward: channel_modulate_bwd (da, db, df in one pass over dout and f), C, B, us, GB/s, %% of measured HBM peak (%.0f)" % peak_gb)
for C in (2048,):
    for B in (16, 256, 1024, 4096):
        f = torch.rand(B, V, C, device="cuda"); g = torch.randn(B, V, C, device="cuda"); a = torch.randn(B, C, device="cuda")
        for want_df in (False, True):
            t = timeit(lambda: ops.channel_modulate_bwd(g, f, a, want_df=want_df))
            byt = 4.0 * (2 * B * V * C + 2 * B * C + (B * V * C if want_df else 0))
            print("a2 channel_modulate_bwd df=%d C=%4d B=%5d %9.1f us %8.1f GB/s %5.1f%%" % (want_df, C, B, t * 1e6, byt / t / 1e9,
                                                                                           100 * byt / t / 1e9 / peak_gb), flush=True)
        del f, g, a
