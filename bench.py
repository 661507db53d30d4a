"""Headline benchmark: agent_dg navigation steps/sec on synthetic R2R-shaped features (BASELINE.json configs[1]):
one "step" = one teacher-forced vl_rollout of B=20 episodes x T=35 actions, forward + backward + the RMSprop
optimizer step (agent_dg.py:633-1033, 1389-1405), full geometry (36 views x 2176, 80-token instructions, hidden 1024,
9 language + 3 cross-modal layers recomputed every navigation step exactly like the reference — no caching).

  python bench.py --gpus N --steps K --warmup W            # our arm (one process per GPU under torchrun for N > 1)
  python bench.py --impl reference ...                     # the reference algorithm on the host CPU (oracle port)

Prints ONE JSON line (see the driver contract). value = episodes x actions of all ranks / max-over-ranks device time.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

B_DEFAULT, T_DEFAULT = 20, 35
ML_WEIGHT = 0.4           # --mlWeight_org (README.md:84)
LR = 1e-4


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="dasa_b200", choices=["dasa_b200", "reference"])
    ap.add_argument("--batch", type=int, default=B_DEFAULT)
    ap.add_argument("--actions", type=int, default=T_DEFAULT)
    ap.add_argument("--precision", default=os.environ.get("DASA_PRECISION", "tf32"), choices=["fp32", "tf32"])
    ap.add_argument("--no-graph", action="store_true", help="launch kernels eagerly instead of replaying a CUDA graph")
    ap.add_argument("--skip-cpu-baseline", action="store_true")
    ap.add_argument("--skip-micro", action="store_true")
    ap.add_argument("--skip-env", action="store_true", help="do not time the device-resident-environment end-to-end arm")
    ap.add_argument("--env-viewpoints", type=int, default=512, help="viewpoints of the synthetic navigation graph (e2e_env arm)")
    ap.add_argument("--schedule", default="batched", choices=["batched", "sequential"],
                    help="batched: AdaIN + encoder of all T teacher-forced actions as one batch; sequential: per action")
    ap.add_argument("--skip-sequential", action="store_true", help="do not also time the per-action schedule")
    ap.add_argument("--feedback", default="teacher", choices=["teacher", "sample"],
                    help="teacher: BASELINE configs[1] (default). sample: accumulate_gradient('sample') = teacher-forced rollout + "
                         "sampled A2C rollout per optimizer step (configs[3] with --batch 512)")
    ap.add_argument("--finetune", action="store_true",
                    help="configs[2]: --d_update_add_layer True (gradients through the 3 cross-modal layers + vision encoder), "
                         "batch 2, two accumulate_gradient('sample') passes (GT + augmented env, train.py:226-243) per optimizer "
                         "step, lr 2e-6: the latency-bound small-batch path")
    ap.add_argument("--no-overlap", action="store_true", help="N > 1: blocking all-reduces after the last weight-gradient GEMM")
    ap.add_argument("--skip-config3", action="store_true", help="N > 1: do not also time BASELINE configs[3] (512 episodes/GPU, sampled feedback)")
    ap.add_argument("--skip-config0", action="store_true", help="do not time BASELINE configs[0] (single eval decode step, GPU vs CPU eager)")
    ap.add_argument("--profile-step", action="store_true",
                    help="bracket exactly one eager step with cudaProfilerStart/Stop (ncu --profile-from-start off) and exit")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------ clocks sampler
class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index=0):
        self.samples, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        for s in self.samples:
            f = [x.strip() for x in s.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0]))
                mx = float(f[1])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


# --------------------------------------------------------------------------------------------- CPU reference arm
def cpu_rollout_sample(B, t_sample, threads=None):
    """The reference algorithm (oracle port of the agent_dg modules + loop) on the host CPU: one teacher-forced rollout
    of B episodes truncated to t_sample actions, forward + backward. Returns (seconds, nav steps)."""
    from dasa_b200 import synth
    from dasa_b200.config import FULL
    from oracle import restated as R
    if threads:
        torch.set_num_threads(threads)
    st = synth.policy_state(FULL, 0)
    trainable = ("adaIn", "decoder", "critic")
    for grp in trainable:
        for v in st[grp].values():
            v.requires_grad_(True)
    for k, v in st["encoder"].items():
        if not k.startswith("bert."):
            v.requires_grad_(True)
    ep = synth.Episodes(B, t_sample, FULL, seed=0)
    gen = torch.Generator().manual_seed(0)

    class RandDrops:
        training = True

        def __call__(self, x, p, tag):
            return x * ((torch.rand(x.shape, generator=gen) >= p).to(x.dtype) / (1 - p))
    t0 = time.perf_counter()
    loss, _, _ = R.teacher_rollout(st, FULL, ep, t_sample, ML_WEIGHT, RandDrops())
    loss.backward()
    dt = time.perf_counter() - t0
    return dt, B * t_sample


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    t_sample = 8 if args.batch <= 20 else 1
    for _ in range(max(0, min(args.warmup, 1))):
        cpu_rollout_sample(args.batch, t_sample)
    times, steps = [], 0
    for _ in range(max(1, min(args.steps, 4))):
        dt, n = cpu_rollout_sample(args.batch, t_sample)
        times.append(dt)
        steps += n
    total = sum(times)
    value = steps / total
    sample = "oracle port (CPU torch), B=%d rollout truncated to %d of %d actions per step, fwd+bwd, %d repeats" % (
        args.batch, t_sample, args.actions, len(times))
    print(json.dumps({
        "impl": "reference", "metric": "nav_steps_per_sec", "value": value, "unit": "nav steps/s (episodes x actions)",
        "n_gpus": args.gpus, "steps": len(times), "warmup": min(args.warmup, 1), "ms_per_step": 1e3 * total / len(times),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "agent_dg teacher-forced rollout fwd+bwd, B=%d, T=%d (bounded sample: %d action/step)" % (
            args.batch, args.actions, t_sample)},
        "cpu_baseline": {"value": value, "unit": "nav steps/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "nav steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# ------------------------------------------------------------------------------------------------------ our arm
def micro_rooflines(peak_gbs):
    """Config 5 excerpt: the HBM-bound AdaIN / attention kernels at a large batch against the measured copy bandwidth."""
    from dasa_b200 import ops
    out = {}
    dev = "cuda"
    B, V, C, A = 1024, 36, 2048, 128
    F = C + A
    f = torch.rand(B, V, F, device=dev)
    d = torch.rand(B, V, F, device=dev)
    g = torch.randn(B * V, C, device=dev)
    o = torch.empty(B, V, F, device=dev)
    h_t = torch.randn(B, F, device=dev) * 0.05
    kl = torch.randn(B, 5, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def timeit(fn, reps=6):
        ts = []
        for i in range(reps + 2):
            flush.zero_()                                   # evict L2 between timed launches
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            if i >= 2:
                ts.append(e0.elapsed_time(e1) * 1e-3)
        ts.sort()
        return ts[len(ts) // 2]

    def add(name, bytes_, secs):
        out[name] = {"bound": "hbm", "achieved": bytes_ / secs / 1e9, "peak": peak_gbs, "unit": "GB/s",
                     "frac": bytes_ / secs / 1e9 / peak_gbs, "frac_vs_8tbs_nominal": bytes_ / secs / 1e9 / 8000.0, "batch": B}
    t = timeit(lambda: ops.gate_modulate(g, f[..., :C], o[..., :C]))
    add("adain_gate_modulate", 4 * 3 * B * V * C, t)
    t = timeit(lambda: ops.adain_rows(f[..., :C], d[..., :C], 1e-5, o[..., :C]))
    add("adain_rows(default)", 4 * 3 * B * V * C, t)
    t = timeit(lambda: ops.view_stats(d[..., :C]))
    add("adain_view_stats", 4 * (B * V * C + 4 * B * C), t)
    t = timeit(lambda: ops.row_attention_fwd(f, h_t, None, 5, 12, kl))
    add("shift_attention_fwd", 4 * (B * V * F + 2 * B * F + B * V + B * 5), t)
    # K1 in its GEMM-fused form (what the rollout runs): sigmoid(d W^T + b) * f * keep-mask, gate saved on the side — tensor-bound
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        peaks = {}
    peak_tf = float(peaks.get("bf16_tflops", 1590.0)) / 2.0     # a kernel timed alone: the BURST bf16 figure / 2 (tf32)
    prec = ops.get_precision()
    ops.set_precision("tf32")
    W = torch.randn(C, C, device=dev) / C ** 0.5
    bias = torch.randn(C, device=dev) * 0.02
    R = B * V
    sgate = torch.empty(R, C, device=dev)
    keep = (torch.rand(R, C, device=dev) >= 0.4).to(torch.uint8)
    f2, d2, o2 = f.view(R, F), d.view(R, F), o.view(R, F)
    t = timeit(lambda: ops.gemm(d2, F, 1, W, C, 1, o2, F, R, C, C, epilogue=ops.EPI_GATE, bias=bias, gate_src=f2, ld_gate=F,
                                gate_out=sgate, ld_gate_out=C, drop_mask=keep, drop_scale=1 / 0.6))
    ops.set_precision(prec)
    out["adain_gate_gemm_fused"] = {"bound": "tensor", "achieved": 2.0 * R * C * C / t / 1e12, "peak": peak_tf, "unit": "TFLOP/s",
                                    "frac": 2.0 * R * C * C / t / 1e12 / peak_tf, "batch": B,
                                    "peak_source": "%s bf16_tflops (burst) / 2" % ("measured" if "bf16_tflops" in peaks else "fallback"),
                                    "what": "tcgen05 TF32 GEMM %d x %d x %d with the sigmoid-gate epilogue (f strided in place, keep mask, gate "
                                            "saved)" % (R, C, C)}
    del g, o, d, sgate, keep, W
    B4 = 4096                                               # the large-batch plateau of the persistent pipelined kernel
    f4 = torch.rand(B4, V, F, device=dev)
    h4 = torch.randn(B4, F, device=dev) * 0.05
    kl4 = torch.randn(B4, 5, device=dev)
    t = timeit(lambda: ops.row_attention_fwd(f4, h4, None, 5, 12, kl4))
    add("shift_attention_fwd@4096", 4 * (B4 * V * F + 2 * B4 * F + B4 * V + B4 * 5), t)
    out["shift_attention_fwd@4096"]["batch"] = B4
    # fused K1 epilogue -> K3: raw features + gate pre-activations in, df_t never materialised
    g4 = torch.randn(B4, V, C, device=dev)
    t = timeit(lambda: ops.gate_shift_attention_fwd(f4, g4, h4, kl4, 5, 12))
    add("fused_gate_shift_attention_fwd@4096", 4 * (B4 * V * (2 * C + A) + 2 * B4 * F + B4 * V), t)
    out["fused_gate_shift_attention_fwd@4096"]["batch"] = B4
    o4 = torch.empty(B4, V, F, device=dev)
    o4[..., C:] = f4[..., C:]
    t_unfused = timeit(lambda: (ops.gate_modulate(g4.view(B4 * V, C), f4[..., :C], o4[..., :C]),
                                ops.row_attention_fwd(o4, h4, None, 5, 12, kl4)))
    out["fused_gate_shift_attention_fwd@4096"]["speedup_vs_unfused_pair"] = t_unfused / t
    return out


def bind_to_gpu_numa_node(local):
    """Pin this rank's threads to the CPUs of its GPU's NUMA node BEFORE any pinned host buffer is allocated (first touch puts
    the pages on that node): at N = 8 every rank uploads 573 MB per step and remote-socket pinned memory shares one UPI link.
    Returns a short description for the JSON line; never fatal."""
    try:
        import pynvml
        pynvml.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        idx = int(vis.split(",")[local]) if vis and all(x.strip().isdigit() for x in vis.split(",")) else local
        h = pynvml.nvmlDeviceGetHandleByIndex(idx)
        pynvml.nvmlDeviceSetCpuAffinity(h)
        return "cpus %d (nvmlDeviceSetCpuAffinity)" % len(os.sched_getaffinity(0))
    except Exception as e:
        return "unbound (%s)" % repr(e)[:80]


def run_ours(args):
    import torch.distributed as dist
    from dasa_b200 import lib, modules as M, ops, synth
    from dasa_b200.config import FULL
    from dasa_b200.rollout import DeviceEpisodes, NavPolicy
    from dasa_b200.trainer import RolloutTrainer
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = "cuda:%d" % local
    world_env = int(os.environ.get("WORLD_SIZE", "1"))
    numa = bind_to_gpu_numa_node(local) if world_env > 1 else "unbound (single rank: the cpu_baseline leg uses every host core)"
    if world > 1:
        import datetime
        dist.init_process_group("nccl", device_id=torch.device(dev), timeout=datetime.timedelta(seconds=600))
    lib.load()
    ops.set_precision(args.precision)
    from dasa_b200 import functions as Fn
    Fn.defer_weight_grads(args.batch <= 64)      # one long-K weight-gradient GEMM per weight per rollout (holds dY, X until then)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak_gbs = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured" if "hbm_gbs" in peaks else "fallback"

    if args.finetune:
        from dataclasses import replace
        args.feedback = "sample"
        if args.batch == B_DEFAULT:
            args.batch = 2
    cfg, B, T = (replace(FULL, update_add_layer=True) if args.finetune else FULL), args.batch, args.actions
    n_acc = 2 if args.finetune else 1                       # GT env + augmented env per optimizer step
    ml_weights = [ML_WEIGHT, 0.6][:n_acc]                   # --mlWeight_org / --mlWeight_aug (param.py:58-59)
    lr = 2e-6 if args.finetune else LR
    pol = NavPolicy(cfg, synth.policy_state(cfg, 0), dev).train()
    pol.schedule = args.schedule
    sample = args.feedback == "sample"
    if sample:
        args.schedule = "sequential"                        # the next observation depends on the sampled action
        pol.schedule = "sequential"
    host_ep = synth.Episodes(B, T + (1 if sample else 0), cfg, seed=100 + rank, pin=True)
    ep_res = DeviceEpisodes(host_ep, dev, resident=True)
    src = M.DropoutSource(seed=1234 + rank, device_seed=True, device=dev)
    loss_host = torch.zeros(1).pin_memory()
    # the product API for one optimizer step (dasa_b200/trainer.py): per-rank loss scaling / global A2C normaliser, deferred
    # weight-gradient GEMMs flushed group by group with each group's NCCL all-reduce overlapped, clip + RMSprop, and the whole
    # thing captured as ONE CUDA graph at any world size
    tr = RolloutTrainer(pol, T, feedback=args.feedback, lr=lr, world=world, dropout_source=src, passes=ml_weights,
                        overlap=not args.no_overlap)
    tr.broadcast_parameters()
    state = {"graph_error": None}

    # e2e: every rollout's inputs come from pinned host memory. The copy of rollout i+1 runs on a side stream into a staging
    # set while rollout i computes (one device-to-device hand-over per step).
    up = {"stream": None, "stage": None, "ready": None, "free": None, "pending": False}

    def start_upload():
        if up["stream"] is None:
            up["stream"] = torch.cuda.Stream()
            up["stage"] = {k: torch.empty_like(getattr(ep_res, k)) for k in DeviceEpisodes.FIELDS}
            up["ready"], up["free"] = torch.cuda.Event(), torch.cuda.Event()
            up["free"].record()
        up["stream"].wait_event(up["free"])                 # the staging set has been handed over
        with torch.cuda.stream(up["stream"]):
            for k in DeviceEpisodes.FIELDS:
                up["stage"][k].copy_(getattr(host_ep, k), non_blocking=True)
            up["ready"].record()
        up["pending"] = True

    def one_step(ep, read_back, upload=False, prefetch_next=False):
        if upload:
            if not up["pending"]:
                start_upload()
            torch.cuda.current_stream().wait_event(up["ready"])
            for k in DeviceEpisodes.FIELDS:
                getattr(ep_res, k).copy_(up["stage"][k], non_blocking=True)
            up["free"].record()
            up["pending"] = False
            if prefetch_next:
                start_upload()
        loss = tr.step(ep_res)
        if read_back:
            loss_host.copy_(loss.detach(), non_blocking=True)
        return loss

    def timed(read_back, upload, steps, warmup):
        # e2e (upload=True) is the steady state of the double-buffered pipeline: every step - warm-up steps included - starts the
        # upload of its successor's inputs on the copy stream before it computes, so each timed step consumes an upload started one
        # step earlier; the timed region starts `steps` uploads of its own and closes only after the last of them has landed.
        for _ in range(warmup):
            one_step(ep_res, read_back, upload, prefetch_next=upload)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        l0 = lib.launches
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            one_step(ep_res, read_back, upload, prefetch_next=upload)
        if upload:
            torch.cuda.current_stream().wait_event(up["ready"])
        e1.record()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        n = (lib.launches - l0) // steps
        return float(ms) / steps, (n + tr.graph_launches if tr.graph is not None else n)

    for _ in range(2):                                      # eager warm-up (allocator, caches, smem attributes, NCCL communicator)
        one_step(ep_res, False)
    torch.cuda.synchronize()

    def try_capture(trainer, ep):
        if args.no_graph or args.profile_step:
            return
        if sample:
            # the closed-loop sampled rollout draws from torch's generator and is launched eagerly; a FAILED capture attempt would
            # also leave that generator in its capture state ("Offset increment outside graph capture") for the eager steps
            state["graph_error"] = "sampled feedback: closed-loop rollout, eager launches"
            return
        try:
            trainer.capture(ep)
        except Exception as e:                              # stay on eager launches, say so in the JSON line
            trainer.release_graph()
            state["graph_error"] = repr(e)[:200]
            torch.cuda.synchronize()
    try_capture(tr, ep_res)

    if args.profile_step:
        Fn.invalidate_weight_caches()
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
        one_step(ep_res, False)
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        if world > 1:
            dist.destroy_process_group()
        return
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms_step, launches = timed(False, False, args.steps, args.warmup)
    clocks = sampler.stop() if rank == 0 else None
    launch_mode = ("cuda-graph replay of the whole optimizer step (rollout fwd+bwd, %sclip + RMSprop)" % (
        "NCCL gradient all-reduces, " if world > 1 else "")) if tr.graph is not None else \
        "eager launches (%s)" % (state["graph_error"] or "--no-graph")
    ms_e2e, _ = timed(True, True, args.steps, 1)
    Fn.invalidate_weight_caches()
    env_arm = None
    if not sample and not args.skip_env:
        try:
            env_arm = env_e2e(args, cfg, B, T, dev, world, rank, tr, try_capture, loss_host, host_ep)
        except Exception as e:                              # reported, never fatal for the headline line
            env_arm = {"error": repr(e)[:300]}
        Fn.invalidate_weight_caches()
    # the same workload on the per-action schedule (the order a sampled / greedy rollout is forced to use)
    ms_seq = None
    if args.schedule == "batched" and not args.skip_sequential and not sample:
        tr.release_graph()
        pol.schedule = "sequential"
        torch.cuda.synchronize()
        try:
            for _ in range(2):
                one_step(ep_res, False)
            torch.cuda.synchronize()
            try_capture(tr, ep_res)
            ms_seq, _ = timed(False, False, max(1, args.steps // 2), 1)
        except Exception as e:
            ms_seq = None
            state["graph_error"] = repr(e)[:200]
        tr.release_graph()
        pol.schedule = args.schedule

    # standalone cost of the gradient all-reduce (the four flat buffers, back to back, nothing to overlap with): what the
    # overlapped, captured reduction has to hide
    allreduce_ms = None
    if world > 1:
        tr.release_graph()
        torch.cuda.synchronize()
        dist.barrier()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for i in range(3):
            if i == 1:
                a0.record()
            for buf in pol.grad_buffers():
                dist.all_reduce(buf)
        a1.record()
        torch.cuda.synchronize()
        ar = torch.tensor([a0.elapsed_time(a1) / 2.0], device=dev)
        dist.all_reduce(ar, op=dist.ReduceOp.MAX)
        allreduce_ms = float(ar)
    config3 = None
    if world > 1 and not sample and not args.finetune and not args.skip_config3:
        config3 = config3_arm(args, cfg, T, dev, world, rank, pol, tr, peaks)
    nav = B * T * world * (2 if sample else 1) * n_acc    # sample feedback: a teacher-forced and a sampled rollout per pass; n_acc passes
    value = nav / (ms_step * 1e-3)
    e2e_value = nav / (ms_e2e * 1e-3)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # dominant kernel of the step: the tcgen05/FFMA GEMM family streams the policy's weights; timed live with CUDA events.
    # The captured graphs (and their private memory pools) are released first: at 512 episodes/GPU a graph pool and an eager
    # rollout do not fit the 180 GB together.
    import gc
    tr.release_graph()
    gc.collect()
    torch.cuda.synchronize()
    torch.cuda.empty_cache()
    roof = dominant_kernel_roofline(pol, ep_res, src, peaks, peak_src)
    config0 = None
    if not args.skip_config0 and not sample:
        try:
            config0 = config0_arm(args, cfg, dev, pol, ep_res, host_ep, not args.skip_cpu_baseline and world == 1)
        except Exception as e:
            config0 = {"error": repr(e)[:300]}
    extra = {} if args.skip_micro else micro_rooflines(peak_gbs)
    cpu = None
    if not args.skip_cpu_baseline and world == 1:
        cores = os.cpu_count() or 1
        torch.set_num_threads(cores)
        t_cpu = min(T, 35 if B <= 20 else 2)
        dt, n = cpu_rollout_sample(B, t_cpu)
        cpu = {"value": n / dt, "unit": "nav steps/s", "cores": cores, "kind": "port",
               "sample": "oracle port on host CPU (torch %d threads): one teacher-forced rollout, B=%d, %d of %d actions, fwd+bwd "
                         "(%.1f s)" % (cores, B, t_cpu, T, dt)}
    line = {
        "metric": "nav_steps_per_sec", "value": value, "unit": "nav steps/s (episodes x actions)", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32 (tf32 tensor-core products in the dense projections)" if args.precision == "tf32" else "f32",
        "data": "synthetic",
        "config": {"workload": ("agent_dg teacher-forced vl_rollout fwd+bwd+RMSprop, B=%d/GPU, T=%d, 36x2176 views, 80-token "
                                "instructions, 9 la + 3 vl layers evaluated for every action (the instruction-only la stack of the T actions batched "
                                "into one pass, own dropout masks per action; nothing cached) (BASELINE.json configs[1])" % (B, T)) if not sample else
                               ("FINETUNE (d_update_add_layer: cross-modal layers + vision encoder trained, lr 2e-6, 2 accumulate_gradient "
                                "passes per optimizer step; BASELINE.json configs[2]) - " if args.finetune else "") +
                               ("agent_dg accumulate_gradient('sample'): teacher-forced vl_rollout + sampled-feedback A2C vl_rollout (Categorical "
                                "sampling on the device, critic, A2C epilogue), fwd+bwd+RMSprop, B=%d episodes/GPU, T=%d, synthetic observation "
                                "stream%s" % (B, T, "" if args.finetune else " (BASELINE.json configs[3])")),
                   "precision": args.precision, "l2": "inputs+weights+activations per step (~1 GB) exceed the 126 MB L2",
                   "parallelism": "dp%d" % world,
                   "schedule": ("batched: the agent follows the teacher, so all T observations are known up front and AdaIN + "
                                "cross-modal layers + bi-LSTM of the T actions run as one batch; decoder sequential"
                                if args.schedule == "batched" else "sequential: AdaIN -> encoder -> decoder per action"),
                   "launch": launch_mode},
        "clocks": clocks, "gpu_launches": launches,
        "e2e": {"value": e2e_value, "unit": "nav steps/s", "h2d_bytes_per_step": host_ep.h2d_bytes_per_step() * T,
                "d2h_bytes_per_step": 4},
        "e2e_env": env_arm,
        "host_numa_binding": numa,
        "roofline": roof, "kernels": extra, "cpu_baseline": cpu,
        "sequential_schedule": None if ms_seq is None else {"value": nav / (ms_seq * 1e-3), "unit": "nav steps/s",
                                                             "ms_per_step": ms_seq},
        "graph_ms": ms_step if launch_mode.startswith("cuda-graph") else None, "allreduce_ms": allreduce_ms,
        "allreduce": None if world == 1 else ("4 flat gradient buffers (%.0f MB fp32), each all-reduced asynchronously as soon as its "
                                              "weight-gradient GEMMs are enqueued, largest first; standalone cost allreduce_ms" % (
                                                  sum(b.numel() for b in pol.grad_buffers()) * 4 / 1e6)),
        "config0_single_step": config0, "config3": config3,
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def env_e2e(args, cfg, B, T, dev, world, rank, tr_main, try_capture, loss_host, host_ep):
    """End-to-end arm over the device-resident environment (dasa_b200/env.py): the RGB / depth feature
    banks and the navigation-graph tables live in HBM (the reference keeps them in host RAM and re-uploads every observation),
    so a rollout's host inputs are the episode descriptors (start viewpoint, heading, goal: 12 B per episode) and the tokenised
    instructions. Per step, inside the timed region: H2D of those from pinned memory, T x (observe, step) kernels that unroll the
    teacher-forced trajectories and assemble every observation, the same forward + backward + optimizer work as `value`, and
    the loss read back. New start / goal pairs every step (the trajectories differ; shapes do not, so one CUDA graph replays)."""
    import torch.distributed as dist
    from dasa_b200 import lib, synth
    from dasa_b200.env import DeviceEnv
    from dasa_b200.navgraph import NavGraph
    n = args.env_viewpoints
    g = NavGraph.synthetic(n, seed=0)
    gen = torch.Generator().manual_seed(4242)
    rgb = synth.resnet_like((n, cfg.views, cfg.rgb_size), gen)
    dep = synth.resnet_like((n, cfg.views, cfg.rgb_size), gen)
    env = DeviceEnv(g, rgb, dep, cfg, dev)
    del rgb, dep
    K = 8
    import numpy as np
    sets = [g.sample_episodes(B, seed=1000 * rank + k) for k in range(K)]
    h_start = torch.from_numpy(np.stack([s[0] for s in sets])).pin_memory()
    h_view = torch.from_numpy(np.stack([s[1] for s in sets])).pin_memory()
    h_goal = torch.from_numpy(np.stack([s[2] for s in sets])).pin_memory()
    h_seq, h_mask, h_len = host_ep.seq.pin_memory(), host_ep.seq_mask.pin_memory(), host_ep.seq_lengths.to(torch.int32).pin_memory()
    d_seq, d_mask, d_len = h_seq.to(dev), h_mask.to(dev), h_len.to(dev)
    lengths_host = [int(x) for x in host_ep.seq_lengths.tolist()]
    h2d = 3 * B * 4 + h_seq.numel() * 8 + h_mask.numel() + h_len.numel() * 4

    def make_ep():
        return env.teacher_episodes(T, (d_seq, d_mask, d_len, lengths_host))

    def upload(i):
        d_seq.copy_(h_seq, non_blocking=True)
        d_mask.copy_(h_mask, non_blocking=True)
        d_len.copy_(h_len, non_blocking=True)
        env.reset(h_start[i % K], h_view[i % K], h_goal[i % K])

    from dasa_b200.trainer import RolloutTrainer
    tr = RolloutTrainer(tr_main.pol, T, feedback="teacher", lr=tr_main.lr, world=world, dropout_source=tr_main.src,
                        passes=tr_main.ml_weights, overlap=tr_main.overlap)

    def step(i):
        upload(i)
        loss = tr.step(make_ep) if tr.graph is not None else tr.step_eager(make_ep())
        loss_host.copy_(loss.detach(), non_blocking=True)

    env.reset(h_start[0], h_view[0], h_goal[0])
    for i in range(2):
        step(i)
    torch.cuda.synchronize()
    env.check()
    try_capture(tr, make_ep)
    mode = "cuda-graph replay (environment unroll + rollout + optimizer)" if tr.graph is not None else "eager launches"
    for i in range(2):
        step(i)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    l0 = lib.launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        step(2 + i)
    e1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    env.check()
    ms = float(ms) / args.steps
    n_launch = (lib.launches - l0) // args.steps + (tr.graph_launches if tr.graph is not None else 0)
    tr.release_graph()
    # HBM roofline of the observation gather at a large batch (L2 flushed between launches)
    obs_roof = None
    if rank == 0 and not args.skip_micro:
        Bm = 1024
        sb = g.sample_episodes(Bm, seed=77)
        env_m = DeviceEnv.__new__(DeviceEnv)
        env_m.__dict__.update(env.__dict__)                 # same resident tables / banks, own episode state
        env_m.B = 0
        env_m.reset(*sb)
        bufm = env_m.alloc(1)
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
        ts = []
        for i in range(8):
            flush.zero_()
            a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a0.record()
            env_m.observe(bufm, 0)
            a1.record()
            torch.cuda.synchronize()
            if i >= 2:
                ts.append(a0.elapsed_time(a1) * 1e-3)
        ts.sort()
        # algorithmic bytes: per view / live candidate row read 2 x C floats (RGB + depth bank rows) and write 2 x (C + A);
        # END / padding rows are written only
        live_rows = Bm * cfg.views + int(g.deg[sb[0]].sum())
        all_rows = Bm * (cfg.views + env.nc)
        bytes_ = 4.0 * (live_rows * 2 * cfg.rgb_size + all_rows * 2 * cfg.feat)
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("hbm_gbs", 6650.0)) \
            if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
        t_med = ts[len(ts) // 2]
        obs_roof = {"bound": "hbm", "achieved": bytes_ / t_med / 1e9, "peak": peak, "unit": "GB/s",
                    "frac": bytes_ / t_med / 1e9 / peak, "batch": Bm, "us": t_med * 1e6}
        del bufm, flush
    return {"value": B * T * world / (ms * 1e-3), "unit": "nav steps/s", "ms_per_step": ms, "h2d_bytes_per_step": h2d,
            "d2h_bytes_per_step": 4, "gpu_launches": n_launch, "launch": mode, "env_observe_roofline": obs_roof,
            "what": "feature banks (%d viewpoints x 36 views x 2048 x {RGB, depth} = %.0f MB) + graph tables resident in HBM; per step the "
                    "episode descriptors and tokenised instructions are uploaded, the teacher-forced trajectories are unrolled and "
                    "every observation is assembled on the device (env_observe / env_step kernels), then the same fwd+bwd+RMSprop "
                    "as `value`" % (n, 2 * n * cfg.views * cfg.rgb_size * 4 / 1e6)}


def config0_arm(args, cfg, dev, pol, ep, host_ep, with_cpu):
    """BASELINE.json configs[0]: ONE agent_dg decode step (the loop body of vl_rollout up to the masked logits, agent_dg.py:727-841:
    AdaIN gate on views + candidates, the 9 + 3 layer encoder, bi-LSTM, decoder), eval mode, forward only, batch 20, on the GPU
    (CUDA-graph replay of the step, L2 flushed between replays) next to the reference algorithm on the host CPU (oracle port,
    CPU eager, all cores)."""
    from dasa_b200 import lib
    was_training = pol.decoder.training
    pol.eval()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    out = {}
    try:
        with torch.no_grad():
            for _ in range(2):
                pol.step(ep, 0, None)
            torch.cuda.synchronize()
            l0 = lib.launches
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                logit, h_t, carry = pol.step(ep, 0, None)
            n_launch = lib.launches - l0
            ts = []
            for i in range(7):
                flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                g.replay()
                e1.record()
                torch.cuda.synchronize()
                if i >= 2:
                    ts.append(e0.elapsed_time(e1))
            ts.sort()
            ms_graph = ts[len(ts) // 2]
            te = []
            for i in range(5):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                pol.step(ep, 0, None)
                e1.record()
                torch.cuda.synchronize()
                te.append(e0.elapsed_time(e1))
            te.sort()
            del g
        out = {"what": "one decode step (AdaIN gates, 9 la + 3 vl layers, bi-LSTM, decoder, masked candidate logits), eval forward, "
                       "B=%d, 36 views x 2176, hidden 1024 (BASELINE.json configs[0])" % ep.B,
               "gpu_ms": ms_graph, "gpu_ms_eager_launches": te[len(te) // 2], "gpu_launches": n_launch,
               "gpu_steps_per_sec": ep.B / (ms_graph * 1e-3), "unit": "nav steps/s", "l2": "flushed between replays"}
    finally:
        if was_training:
            pol.train()
    if with_cpu:
        from dasa_b200 import synth
        from oracle import restated as R
        cores = os.cpu_count() or 1
        torch.set_num_threads(cores)
        st = synth.policy_state(cfg, 0)
        tc = []
        with torch.no_grad():
            for i in range(4):
                t0 = time.perf_counter()
                R.teacher_rollout(st, cfg, host_ep, 1, ML_WEIGHT, R.NoDrop())
                tc.append(time.perf_counter() - t0)
        cpu_ms = 1e3 * sorted(tc[1:])[len(tc[1:]) // 2]
        out.update({"cpu_ms": cpu_ms, "cpu_cores": cores, "cpu_kind": "port (oracle on CPU torch, eager)",
                    "cpu_steps_per_sec": host_ep.B / (cpu_ms * 1e-3), "speedup": cpu_ms / out["gpu_ms"]})
    return out


def config3_arm(args, cfg, T, dev, world, rank, pol, tr_main, peaks):
    """BASELINE.json configs[3] inside the N > 1 run: 512 episodes per GPU, accumulate_gradient('sample') (teacher-forced + sampled
    A2C rollout per optimizer step, A2C loss normalised by the batch-GLOBAL live-action count), NCCL gradient all-reduce, one
    warm-up + two timed optimizer steps, device time, max over ranks. Also the shift-attention / AdaIN-gate HBM rooflines at
    B = 512 (rank 0). Errors are reported, never fatal for the headline line."""
    import gc
    import torch.distributed as dist
    from dasa_b200 import functions as Fn, modules as M, ops, synth
    from dasa_b200.rollout import DeviceEpisodes
    from dasa_b200.trainer import RolloutTrainer
    B3 = 512
    out = {"batch_per_gpu": B3, "feedback": "sample"}
    tr = None
    try:
        gc.collect()
        torch.cuda.empty_cache()
        Fn.defer_weight_grads(False)                        # 512 x 35 rows of dY / X per weight do not need (or fit) deferral
        pol.schedule = "sequential"
        ep = DeviceEpisodes(synth.Episodes(B3, T + 1, cfg, seed=300 + rank), dev, resident=True)
        tr = RolloutTrainer(pol, T, feedback="sample", lr=tr_main.lr, world=world, dropout_source=tr_main.src,
                            overlap=tr_main.overlap)
        tr.step_eager(ep)
        torch.cuda.synchronize()
        dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(2):
            tr.step_eager(ep)
        e1.record()
        torch.cuda.synchronize()
        dist.barrier()
        ms = torch.tensor([e0.elapsed_time(e1) / 2.0], device=dev)
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        nav = B3 * T * 2 * world
        out.update({"value": nav / (float(ms) * 1e-3), "unit": "nav steps/s", "ms_per_step": float(ms), "n_gpus": world,
                    "steps": 2, "warmup": 1, "launch": "eager launches",
                    "peak_mem_gb": torch.cuda.max_memory_allocated() / 1e9})
        del ep
        if rank == 0:
            peak = float(peaks.get("hbm_gbs", 6650.0))
            V, C, A = cfg.views, cfg.rgb_size, cfg.angle_size
            F = C + A
            f = torch.rand(B3, V, F, device=dev)
            gsrc = torch.randn(B3 * V, C, device=dev)
            o = torch.empty(B3, V, F, device=dev)
            h_t = torch.randn(B3, F, device=dev) * 0.05
            kl = torch.randn(B3, 5, device=dev)
            flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

            def timeit(fn):
                ts = []
                for i in range(8):
                    flush.zero_()
                    a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    a0.record()
                    fn()
                    a1.record()
                    torch.cuda.synchronize()
                    if i >= 2:
                        ts.append(a0.elapsed_time(a1) * 1e-3)
                ts.sort()
                return ts[len(ts) // 2]
            roofs = {}
            for name, bytes_, fn in (
                    ("shift_attention_fwd", 4 * (B3 * V * F + 2 * B3 * F + B3 * V + B3 * 5),
                     lambda: ops.row_attention_fwd(f, h_t, None, 5, 12, kl)),
                    ("adain_gate_modulate", 4 * 3 * B3 * V * C, lambda: ops.gate_modulate(gsrc, f[..., :C], o[..., :C]))):
                t = timeit(fn)
                roofs[name] = {"bound": "hbm", "achieved": bytes_ / t / 1e9, "peak": peak, "frac": bytes_ / t / 1e9 / peak,
                               "frac_of_8TBs": bytes_ / t / 1e9 / 8000.0, "unit": "GB/s", "batch": B3}
            out["kernels_at_b512"] = roofs
    except Exception as e:
        out["error"] = repr(e)[:300]
    finally:
        Fn.defer_weight_grads(args.batch <= 64)
        pol.schedule = args.schedule
        Fn._queue.clear()
        gc.collect()
        torch.cuda.empty_cache()
    return out


def dominant_kernel_roofline(pol, ep, src, peaks, peak_src):
    """Event-time every GEMM launch and the two decoder-rollout launches of one training rollout (eager launches, warm caches).
    Dominant kernel of the step by device time (profiles/r02_ncu_launches_step_*_summary.txt): the persistent CTA-pair tcgen05 GEMM
    gemm_tf32_pair_kernel<256,*> (gemm_tc2.cu) - its kind::f16 instantiations run the 78 forward GEMMs of the frozen transformer
    stack and, on fp16 operand copies, the AdaIN gate, the bi-LSTM input projections and the MN-major weight gradients of both
    (dasa_gemm_f16 / dasa_gemm_f16_mn; the bi-LSTM's recurrent GEMMs run inside dasa_bilstm_packed_fwd / _bwd and are not timed
    one by one); its kind::tf32 instantiations keep the remaining token-major GEMMs (vision projection, decoder weight
    gradients). Tensor roofline per family = algorithmic FLOPs (2*M*N*K per launch) / summed launch duration against the
    measured sustained cuBLAS bf16 rate (fp16 operands) or half of it (tf32). The decoder's M = 20-row projections live inside
    the persistent rollout kernels and are reported as a weight-stream (HBM / L2) figure."""
    from dasa_b200 import lib, modules as M
    events, dec = [], []
    orig = lib.call

    def hooked(name, *a):
        if name not in ("dasa_gemm", "dasa_gemm_f16", "dasa_gemm_f16_mn", "dasa_decoder_rollout_fwd", "dasa_decoder_rollout_bwd"):
            return orig(name, *a)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = orig(name, *a)
        e1.record()
        if name == "dasa_gemm":
            events.append((e0, e1, int(a[2]), int(a[3]), int(a[4]), "tf32"))
        elif name in ("dasa_gemm_f16", "dasa_gemm_f16_mn"):
            events.append((e0, e1, int(a[0]), int(a[1]), int(a[2]), "f16"))
        else:
            dec.append((e0, e1, name))
        return rc
    import dasa_b200.ops as ops_mod
    ops_mod.call = hooked
    T = ep.T if ep.dist is None else ep.T - 1
    try:
        pol.zero_grad()
        src.advance()
        with M.use_dropout_source(src):
            loss, _, _ = pol.teacher_rollout(ep, T, ML_WEIGHT, tag_steps=False)
        pol.backward(loss)
        torch.cuda.synchronize()
    finally:
        ops_mod.call = orig
    timed = [(a.elapsed_time(b) * 1e-3, m, n, k, kind) for a, b, m, n, k, kind in events]
    f16 = [x for x in timed if x[4] == "f16"]
    big = [x for x in timed if x[4] == "tf32" and x[1] >= 2048]
    small = [x for x in timed if x[4] == "tf32" and x[1] < 2048]
    peak16 = float(peaks.get("bf16_tflops_sustained", 1400.0))
    hbm = float(peaks.get("hbm_gbs", 6650.0))

    def fam(rows, peak, what):
        secs = sum(t for t, _, _, _, _ in rows)
        flops = sum(2.0 * m * n * k for _, m, n, k, _ in rows)
        ach = flops / max(secs, 1e-12) / 1e12
        return {"bound": "tensor", "what": what, "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak,
                "launches_timed": len(rows), "device_seconds": secs}
    fam16 = fam(f16, peak16, "kind::f16 instantiations (fp16 operands, fp32 accumulate): forward GEMMs of the frozen language / cross-modal stack, AdaIN gate, "
                "bi-LSTM input projections, MN-major weight gradients of the bi-LSTM and the gate")
    fam32 = fam(big, peak16 / 2.0, "kind::tf32 instantiations, token-major GEMMs with M >= 2048 (vision projection, decoder weight gradients)")
    s_secs = sum(t for t, _, _, _, _ in small)
    s_bytes = sum(4.0 * (n * k + m * k + m * n) for _, m, n, k, _ in small)
    # decoder rollout kernels: per action the fp16 weight stream [linear_in ; linear_shift], [W_ih | W_hh], att.linear_in,
    # att.linear_out (forward: as stored, backward: transposed copies) + the fp32 view / instruction-context tiles
    cfg, B = pol.cfg, ep.B
    H, E, F, D, V = cfg.hidden, cfg.action_emb, cfg.feat, cfg.ctx_dim, cfg.views
    w_bytes = 2.0 * ((F + cfg.shift_kernel) * H + 4 * H * (E + F + H) + D * H + H * (D + H))
    tile_bytes = 4.0 * (B * V * F + sum(ep.seq_lengths_host) * D)         # views + the valid instruction tokens of the B episodes
    dec_out = {}
    for e0, e1, name in dec:
        secs = e0.elapsed_time(e1) * 1e-3
        byt = (w_bytes + tile_bytes * (2.0 if name.endswith("bwd") else 1.0)) * T
        dec_out[name.replace("dasa_", "")] = {"bound": "hbm", "achieved": byt / secs / 1e9, "peak": hbm, "unit": "GB/s",
                                               "frac": byt / secs / 1e9 / hbm, "us_per_action": secs * 1e6 / T,
                                               "bytes_per_action": byt / T}
    # top-level = the family with the larger share of the step
    top, other, top_name, other_name = (fam16, fam32, "f16", "tf32") if fam16["device_seconds"] >= fam32["device_seconds"] else \
        (fam32, fam16, "tf32", "f16")
    out = dict(top)
    out.update({
        "kernel": "gemm_tf32_pair_kernel<256,*> (persistent CTA pair, tcgen05.mma cta_group::2 256x256 tiles, TMA 128B-swizzle operands, "
                  "double-buffered TMEM accumulator): %s family" % top_name,
        # one `ncu --set full` capture of ONE launch of this kernel (profiles/r02_gemm_f16_gelu_ncu_full.txt: M=20300 N=3072 K=768, fp16
        # operands and output, GELU epilogue): dram__bytes_read.sum + dram__bytes_write.sum. Not measured in this run.
        "traffic": 110.8e6, "traffic_launch": "M=20300 N=3072 K=768 fp16 in/out (algorithmic 160.6e6 B, 95.8 GFLOP; the output is still "
                                              "partly L2-resident when the kernel ends)",
        "traffic_source": "profiles/r02_gemm_f16_gelu_ncu_full.txt (ncu --set full, one launch), not measured in this run",
        "peak_source": "%s bf16_tflops_sustained%s" % (peak_src, "" if top_name == "f16" else " / 2 (tf32)"),
        "%s_family" % other_name: other,
        "decoder_rollout": dec_out,
        "small_m_gemms": {"bound": "hbm", "what": "remaining dasa_gemm calls with M < 2048 (critic, init-state linears, skinny shapes)",
                          "achieved": s_bytes / max(s_secs, 1e-12) / 1e9, "peak": hbm, "unit": "GB/s",
                          "frac": s_bytes / max(s_secs, 1e-12) / 1e9 / hbm, "launches_timed": len(small),
                          "device_seconds": s_secs, "note": "eager launches, event-bracketed: includes launch gaps"}})
    return out


def main():
    """Everything that libraries write to stdout while the benchmark runs (e.g. NCCL's version banner) goes to stderr: stdout
    carries exactly ONE line, the JSON record."""
    a = parse()
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    lines = []
    import builtins
    orig_print = builtins.print

    def capture_print(*args, **kw):
        if kw.get("file") in (None, sys.stdout) and len(args) == 1 and isinstance(args[0], str) and args[0].startswith("{"):
            lines.append(args[0])
        else:
            orig_print(*args, **kw)
    builtins.print = capture_print
    try:
        if a.impl == "reference":
            run_reference(a)
        else:
            run_ours(a)
    finally:
        builtins.print = orig_print
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
        os.close(real_stdout)
    for ln in lines[-1:]:
        print(ln, flush=True)


if __name__ == "__main__":
    main()
