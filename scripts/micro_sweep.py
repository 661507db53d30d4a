"""BASELINE.json configs[4]: AdaIN + shift-attention microbenchmark sweep against the HBM roofline.
views 36, channels C in {2048, 3072, 4096} (+128 angle columns for the attention), batch 1..8192, L2 flushed between launches,
CUDA-event timed, median of 7. Algorithmic bytes per SURVEY.md section 8(d). Writes a table to stdout (profiles/ keeps a copy)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dasa_b200 import ops

peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]
dev = "cuda"
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timeit(fn, reps=7):
    ts = []
    for i in range(reps + 2):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        if i >= 2:
            ts.append(e0.elapsed_time(e1) * 1e-3)
    ts.sort()
    return ts[len(ts) // 2]


V, A = 36, 128
batches = [1, 4, 16, 64, 256, 512, 1024, 2048, 4096, 8192]
channels = (2048, 3072, 4096)
only = None
if len(sys.argv) > 1:      # quick runs: micro_sweep.py "2048,4096" "512,4096" ["K3,fused"]
    channels = tuple(int(x) for x in sys.argv[1].split(","))
    if len(sys.argv) > 2:
        batches = [int(x) for x in sys.argv[2].split(",")]
    if len(sys.argv) > 3:
        only = sys.argv[3].split(",")
rows = []
print("# kernel, C, B, us, GB/s, %% of measured HBM peak (%.0f GB/s), %% of the nominal 8 TB/s" % peak)
for C in channels:
    F = C + A
    for B in batches:
        if B * V * F * 4 * 3 > 60e9:
            continue
        f = torch.rand(B, V, F, device=dev); d = torch.rand(B, V, F, device=dev); o = torch.empty(B, V, F, device=dev)
        g = torch.randn(B * V, C, device=dev)
        g3 = g.view(B, V, C)
        o[..., C:] = f[..., C:]
        t_ = torch.randn(B, F, device=dev) * 0.05; kl = torch.randn(B, 5, device=dev)
        cases = [
            ("K1 gate_modulate (sigmoid(g)*f, strided in place)", 4 * 3 * B * V * C, lambda: ops.gate_modulate(g, f[..., :C], o[..., :C])),
            ("K2 adain_rows (default AdaIN, one pass)", 4 * 3 * B * V * C, lambda: ops.adain_rows(f[..., :C], d[..., :C], 1e-5, o[..., :C])),
            ("K2 view_stats (mean/std/max/min over views)", 4 * (B * V * C + 4 * B * C), lambda: ops.view_stats(d[..., :C])),
            ("K3 shift_attention_fwd (k=5)", 4 * (B * V * F + 2 * B * F + B * V + B * 5), lambda: ops.row_attention_fwd(f, t_, None, 5, 12, kl)),
            # fused K1 -> K3: gate epilogue feeding the shift attention, df_t never materialised (4*B*V*(2C + A) + 4*(2*B*F + B*V))
            ("K1->K3 fused gate + shift attention", 4 * (B * V * (2 * C + A) + 2 * B * F + B * V),
             lambda: ops.gate_shift_attention_fwd(f, g3, t_, kl, 5, 12)),
            # the same result unfused: gate_modulate writes df_t, the attention re-reads it (bytes: those two kernels' own)
            ("K1 then K3, unfused (df_t through HBM)", 4 * 3 * B * V * C + 4 * (B * V * F + 2 * B * F + B * V + B * 5),
             lambda: (ops.gate_modulate(g, f[..., :C], o[..., :C]), ops.row_attention_fwd(o, t_, None, 5, 12, kl))),
        ]
        for name, byt, fn in cases:
            if only and not any(o in name for o in only):
                continue
            s = timeit(fn)
            gbs = byt / s / 1e9
            rows.append((name, C, B, s * 1e6, gbs, gbs / peak))
            print("%-52s C=%4d B=%5d %9.1f us %8.1f GB/s %5.1f%% %5.1f%%" % (name, C, B, s * 1e6, gbs, 100 * gbs / peak, gbs / 80.0), flush=True)
        del f, d, o, g
