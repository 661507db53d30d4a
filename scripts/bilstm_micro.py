"""GPU helper: packed bi-LSTM forward / forward+backward (R = 700 sequences, H = 1024) as a CUDA-graph replay, fused cell epilogue +
fp16 recurrence operands vs the two-launch TF32 form. argv: modes (0 / 1), e.g. `bilstm_micro.py 0 1`; BWD=1 adds the backward pass."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dasa_b200 import functions as Fn, modules as M, ops

DEV = "cuda"
R, L, In, H = 700, 80, 768, 1024
BWD = os.environ.get("BWD", "0") == "1"
BENCH_LENS = os.environ.get("BENCH_LENS", "0") == "1"
g = torch.Generator().manual_seed(0)
if BENCH_LENS:      # the bench's distribution: 20 instructions x 35 actions
    lens = [45, 40, 39, 39, 37, 35, 34, 34, 34, 32, 30, 29, 28, 27, 25, 22, 10, 9, 9, 8] * 35
else:
    lens = torch.randint(8, 52, (R,), generator=g).tolist()
pack = M.PackInfo(lens, L, 1, DEV)
plan = pack.bilstm_plan()
x = (torch.randn(pack.ntok, In, generator=g) * 0.5).to(DEV)
ws = [(torch.randn(*s, generator=g) * sc).to(DEV).requires_grad_(BWD) for s, sc in
      (((4 * H, In), In ** -0.5), ((4 * H, H), H ** -0.5), ((4 * H,), 0.1), ((4 * H,), 0.1)) * 2]
gout = torch.randn(R, L, 2 * H, generator=g).to(DEV) * 0.01
ops.set_precision("tf32")
Fn.defer_weight_grads(False)
modes = [int(a) for a in sys.argv[1:]] or [0, 1]


def run():
    if not BWD:
        with torch.no_grad():
            return Fn.PackedBiLSTMFn.apply(x, plan, *ws)
    for w in ws:
        w.grad = None
    out, h, c = Fn.PackedBiLSTMFn.apply(x, plan, *ws)
    (out * gout).sum().backward()


for fused in modes:
    ops.fused_lstm_cell = bool(fused)
    for _ in range(2):
        run()
    torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        run()
    gr.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        gr.replay()
    e1.record(); torch.cuda.synchronize()
    print("fused=%d bwd=%d tokens=%d steps=%d  %.3f ms (incl. input projections%s)" %
          (fused, BWD, pack.ntok, max(lens), e0.elapsed_time(e1) / 10, ", weight gradients" if BWD else ""))
