"""GPU helper: packed bi-LSTM forward (R = 700 sequences, H = 1024) as a CUDA-graph replay, fused cell epilogue vs two-launch form."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dasa_b200 import functions as Fn, modules as M, ops

DEV = "cuda"
R, L, In, H = 700, 80, 768, 1024
g = torch.Generator().manual_seed(0)
lens = torch.randint(8, 52, (R,), generator=g).tolist()
pack = M.PackInfo(lens, L, 1, DEV)
plan = pack.bilstm_plan()
x = (torch.randn(pack.ntok, In, generator=g) * 0.5).to(DEV)
ws = [(torch.randn(*s, generator=g) * sc).to(DEV) for s, sc in
      (((4 * H, In), In ** -0.5), ((4 * H, H), H ** -0.5), ((4 * H,), 0.1), ((4 * H,), 0.1)) * 2]
ops.set_precision("tf32")
modes = [int(a) for a in sys.argv[1:]] or [0, 1]
for fused in modes:
    ops.fused_lstm_cell = bool(fused)
    with torch.no_grad():
        for _ in range(2):
            out = Fn.PackedBiLSTMFn.apply(x, plan, *ws)
        torch.cuda.synchronize()
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr):
            out = Fn.PackedBiLSTMFn.apply(x, plan, *ws)
        gr.replay(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            gr.replay()
        e1.record(); torch.cuda.synchronize()
    print("fused=%d  tokens=%d  steps=%d  %.3f ms per forward (incl. input projections)" % (fused, pack.ntok, max(lens), e0.elapsed_time(e1) / 10))
