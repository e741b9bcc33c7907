#!/usr/bin/env python
"""bench.py -- positive-pairs/sec of cl-ica's InfoNCE training step on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c2|c3]
                    [--scaling weak|strong] [--gemm-mode 3xtf32|tf32|fp32]

A "step" is the body of main_mlp.py:258-285 (unsupervised branch): zero_grad -> h(z1) -> h(z2) -> roll ->
LpSimCLRLoss -> backward -> Adam.step, h = f o g with f the 7-layer MLP encoder and g a frozen 3-layer mixing
net, on synthetic latents of the named shape (c2: n=10, B=6144 per GPU, sphere, p=2, tau=1; c3: n=40,
B=8192, p=3).  One JSON line is printed by rank 0 (see DESIGN.md "Measurement" for every key).

  value   device-timed pairs/s with inputs resident in HBM (CUDA events, max over ranks)
  e2e     same step through the reference-facing drop-in modules with HOST buffers: pinned z1/z2 are
          copied host->device inside the timed region every step and the loss + 2 parts are read back
          (.item(), as main_mlp.py:285 does)
  --impl reference: the reference's CPU formulation (oracle/torch_port.py, "port") timed on this box's host
          cores on a bounded sample of the same workload (see oracle.torch_port.sampled_train_step).
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: n, per-GPU batch (weak) / global batch (strong), p, tau, space
    "c2": dict(n=10, B=6144, p=2, tau=1.0, space="sphere",
               desc="main_mlp.py --n 10 --space-type sphere --p 2 --tau 1.0 --batch-size 6144"),
    "c3": dict(n=40, B=8192, p=3, tau=1.0, space="real",
               desc="main_mlp.py --n 40 --space-type unbounded --m-p 2 --c-p 3 --p 3 --batch-size 8192"),
}
FAMILIES = ["loss_fwd", "loss_bwd", "loss_aux", "gemm_tc", "gemm_simt", "adam", "misc"]


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--gemm-mode", default=os.environ.get("CLICA_GEMM_MODE", "3xtf32"))
    ap.add_argument("--cpu-sample-rows", type=int, default=0, help="anchors per CPU-baseline step (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", dest="graph", action="store_false",
                    help="time the eager step instead of the CUDA-graph replay (single GPU)")
    ap.set_defaults(graph=os.environ.get("CLICA_GRAPH", "1") != "0")
    ap.add_argument("--graph-multi", dest="graph_multi", action="store_true",
                    help="multi-GPU: record the sharded step (incl. its NCCL collectives) into the CUDA graph too")
    ap.add_argument("--no-graph-multi", dest="graph_multi", action="store_false")
    ap.set_defaults(graph_multi=os.environ.get("CLICA_GRAPH_MULTI", "1") != "0")
    return ap.parse_args()


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            pk = json.load(fh)
        return dict(hbm_gbs=pk["hbm_gbs"], bf16_tflops=pk["bf16_tflops"],
                    bf16_tflops_sustained=pk.get("bf16_tflops_sustained", pk["bf16_tflops"]), source="measured")
    return dict(hbm_gbs=6650.0, bf16_tflops=1590.0, bf16_tflops_sustained=1400.0, source="fallback")


# ---------------------------------------------------------------------------------------- CPU baseline ----
def cpu_baseline_run(wl, steps, warmup, rows=0, budget_s=25.0):
    """Times oracle.torch_port (the reference's torch formulation) on the host cores on a bounded sample."""
    import torch
    from oracle import torch_port as tp
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    n, B, p, tau = wl["n"], wl["B"], wl["p"], wl["tau"]
    torch.manual_seed(0)
    f = tp.build_encoder(n)
    g = tp.build_mixing(n, 3, seed=0)
    opt = torch.optim.Adam(f.parameters(), lr=1e-4)
    z1, z2 = tp.synth_latents(B, n, wl["space"], seed=0)
    with torch.no_grad():
        z3_rec = f(g(z1)).detach().clone().requires_grad_(True)      # all B negatives, pre-encoded
    if rows <= 0:
        # calibrate: one tiny step, then size the sample so that (warmup + steps) fit the budget
        r0 = min(128, B)
        t0 = time.perf_counter()
        tp.sampled_train_step(f, g, opt, z1, z2, z3_rec, r0, p, tau)
        per_row = (time.perf_counter() - t0) / r0
        rows = int(budget_s / max(per_row * (steps + warmup), 1e-9))
        rows = max(64, min(B, rows // 64 * 64))
    for _ in range(warmup):
        tp.sampled_train_step(f, g, opt, z1, z2, z3_rec, rows, p, tau)
    t0 = time.perf_counter()
    for _ in range(steps):
        tp.sampled_train_step(f, g, opt, z1, z2, z3_rec, rows, p, tau)
    dt = (time.perf_counter() - t0) / steps
    return dict(value=rows / dt, unit="pairs/s", cores=cores, kind="port",
                sample=f"{rows} of {B} anchor/positive pairs per step (encoder fwd+bwd on the sample, loss "
                       f"fwd+bwd against all {B} negatives, Adam); {steps} steps, torch {torch.__version__} CPU",
                ms_per_step=dt * 1e3, rows=rows)


def torch_eager_gpu_run(wl, dev, steps=5, warmup=2):
    """The reference's torch formulation (oracle.torch_port.train_step: materialised B x B x d, cuBLAS fp32 encoder,
    torch.optim.Adam) on the SAME GPU -- the kernel-to-beat baseline of SURVEY.md 8(d).  Baseline only; a failure
    (e.g. out of memory at a large sweep point) is reported as null."""
    import torch
    from oracle import torch_port as tp
    n, B, p, tau = wl["n"], wl["B"], wl["p"], wl["tau"]
    prev = torch.backends.cuda.matmul.allow_tf32
    try:
        torch.backends.cuda.matmul.allow_tf32 = False          # the reference's default: true fp32 matmuls
        torch.manual_seed(0)
        f = tp.build_encoder(n).to(dev)
        g = tp.build_mixing(n, 3, seed=0).to(dev)
        opt = torch.optim.Adam(f.parameters(), lr=1e-4)
        z1, z2 = tp.synth_latents(B, n, wl["space"], seed=0)
        z1, z2 = z1.to(dev), z2.to(dev)
        for _ in range(warmup):
            tp.train_step(f, g, opt, z1, z2, p, tau)
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        for _ in range(steps):
            tp.train_step(f, g, opt, z1, z2, p, tau)         # ends with .item() like main_mlp.py:285
        torch.cuda.synchronize(dev)
        dt = (time.perf_counter() - t0) / steps
        peak_gb = torch.cuda.max_memory_allocated(dev) / 1e9
        return dict(value=B / dt, unit="pairs/s", ms_per_step=dt * 1e3, steps=steps, peak_mem_gb=peak_gb,
                    what="oracle/torch_port.train_step = the reference's eager torch formulation on this GPU "
                         "(fp32 cuBLAS, TF32 off, B x B x d materialised), full batch")
    except Exception as exc:                                   # baseline only
        return dict(value=None, error=repr(exc)[:200])
    finally:
        torch.backends.cuda.matmul.allow_tf32 = prev
        try:
            del f, g, opt, z1, z2
        except Exception:
            pass
        torch.cuda.empty_cache()


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    wl = WORKLOADS[args.workload]
    res = cpu_baseline_run(wl, max(1, args.steps), max(0, args.warmup), args.cpu_sample_rows, budget_s=120.0)
    line = {
        "impl": "reference", "metric": "positive-pairs/sec", "value": res["value"], "unit": "pairs/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": res["ms_per_step"],
        "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        # same workload keys as the CUDA arm's line (the sample is described in cpu_baseline.sample)
        "config": {"workload": wl["desc"], "n": wl["n"], "batch_per_gpu": wl["B"], "global_batch": wl["B"],
                   "p": wl["p"], "tau": wl["tau"], "parallelism": "host CPU threads (torch intra-op)"},
        "cpu_baseline": {k: res[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": res["value"], "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------- clocks ----
class ClockSampler:
    """nvidia-smi sampling DURING the timed region (B200_PROFILING.md's clocks line)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu_index = gpu_index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.fh = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.gpu_index)], stdout=self.fh,
                                         stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        try:
            self.proc.terminate()
            self.proc.wait(timeout=5)
            self.fh.close()
            sm, mx, reasons = [], [], set()
            for ln in open(self.path):
                parts = [x.strip() for x in ln.split(",")]
                if len(parts) < 9:
                    continue
                try:
                    sm.append(float(parts[1])), mx.append(float(parts[2]))
                except ValueError:
                    continue
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), parts[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            if sm:
                sm.sort()
                out.update(sm_mhz=sm[len(sm) // 2], sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm))
        except Exception:
            pass
        finally:
            try:
                os.unlink(self.path)
            except Exception:
                pass
        return out


# ------------------------------------------------------------------------------------------- our arm ----
def algorithmic_work(wl, B_local, B_global, world):
    """Per-step, per-rank algorithmic work: encoder flops and loss-kernel lane-ops / bytes."""
    n, p = wl["n"], wl["p"]
    widths = [n, 10 * n, 50 * n, 50 * n, 50 * n, 50 * n, 10 * n, n]
    mac_row = sum(widths[i] * widths[i + 1] for i in range(7))
    enc_flops = 2 * B_local * (6 * mac_row - 2 * widths[0] * widths[1])       # fwd + dX + dW, 2 calls, no layer-0 dX
    cp = {1: 2, 2: 2, 3: 3}.get(p, 3)
    pe = B_local * B_global * n
    loss_ops_fwd = cp * pe
    loss_ops_bwd = (cp + 2) * pe
    loss_bytes = 24 * B_local * n + 8 * B_local if world == 1 else (12 * B_local * n + 4 * B_global * n + 8 * B_local)
    return dict(enc_flops=enc_flops, loss_ops_fwd=loss_ops_fwd, loss_ops_bwd=loss_ops_bwd, loss_bytes=loss_bytes,
                mac_row=mac_row)


def run_ours(args):
    import torch
    import torch.distributed as dist
    import clica_b200
    from clica_b200 import _lib, sharded
    from clica_b200.optim import FusedAdam
    from clica_b200 import synth as tp

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch multi-GPU runs with torch.distributed.run (see the module docstring)")
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback for the product path)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    lib = _lib.load()
    os.environ["CLICA_GEMM_MODE"] = args.gemm_mode
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    wl = WORKLOADS[args.workload]
    n, p, tau = wl["n"], wl["p"], wl["tau"]
    if args.scaling == "weak":
        B_local, B_global = wl["B"], wl["B"] * world
    else:
        assert wl["B"] % world == 0
        B_local, B_global = wl["B"] // world, wl["B"]

    sys.path.insert(0, clica_b200.DROPIN_DIR)
    import encoders
    import losses

    torch.manual_seed(0)                                   # identical initial weights on every rank
    f = encoders.get_mlp(n, n, [10 * n, 50 * n, 50 * n, 50 * n, 50 * n, 10 * n]).to(dev)
    g = tp.build_mixing(n, 3, seed=0).to(dev)
    crit = losses.LpSimCLRLoss(p=p, tau=tau, simclr_compatibility_mode=True)
    z1_h, z2_h = tp.synth_latents(B_global, n, wl["space"], seed=0)
    z1_h = z1_h[rank * B_local:(rank + 1) * B_local].contiguous().pin_memory()
    z2_h = z2_h[rank * B_local:(rank + 1) * B_local].contiguous().pin_memory()
    z1_d, z2_d = z1_h.to(dev), z2_h.to(dev)
    h = lambda z: f(g(z))

    # ---- the step, device-resident flavour (value): fused Adam, no host sync inside the loop ------------
    opt = FusedAdam(f.parameters(), lr=1e-4)

    # anchors and positives go through the encoder as ONE 2B-row batch (same arithmetic per row; the GEMMs get
    # twice the rows per launch); the e2e flavour below keeps the script's two separate calls
    z12_d = torch.cat([z1_d, z2_d], 0).contiguous()

    def step_eager():
        if world == 1:
            opt.zero_grad(set_to_none=True)
            ab = h(z12_d)
            a, b = ab[:B_local], ab[B_local:]
            total, _, parts = crit(None, None, None, a, b, torch.roll(a, 1, 0))
            total.backward()
            opt.step()
            return total
        total, parts = sharded.sharded_train_step(f, g, opt, z1_d, z2_d, p, tau, 0.5, z12_local=z12_d)
        return total

    # single GPU: the same step recorded once into a CUDA graph (clica_b200.graphed.GraphedTrainStep) and
    # replayed -- no host work between the ~45 kernels of a step.  CLICA_GRAPH=0 times the eager step instead.
    step_mode = "eager"
    graphed = None
    watchdog = None
    if args.graph and (world == 1 or args.graph_multi):
        from clica_b200.graphed import GraphedTrainStep
        if world > 1:
            # recording NCCL collectives into a graph cannot be abandoned in-process: if it wedges, fail fast and
            # loudly instead of sitting in the launcher's timeout (CLICA_GRAPH_MULTI=0 selects the eager sharded step)
            import threading

            def _wedged():
                sys.stderr.write("bench.py: multi-GPU CUDA-graph capture did not finish within 150 s; "
                                 "rerun with CLICA_GRAPH_MULTI=0\n")
                sys.stderr.flush()
                os._exit(17)
            watchdog = threading.Timer(150.0, _wedged)
            watchdog.daemon = True
            watchdog.start()
        try:
            graphed = GraphedTrainStep(f, g, crit, B_local, n, lr=1e-4, host_io=False,
                                       group=(dist.group.WORLD if world > 1 else None))
            graphed.stage(z1_d, z2_d)
            graphed.replay()
            torch.cuda.synchronize()
            step_mode = "cuda_graph"
        except Exception as exc:           # same kernels either way: time the eager step and say so in the JSON line
            if world > 1:
                raise                      # ranks must not diverge on the step flavour
            sys.stderr.write(f"bench.py: CUDA-graph step unavailable ({exc!r}); timing the eager step\n")
            graphed = None
            args.graph = False
            step_mode = "eager (graph capture failed: " + repr(exc)[:120] + ")"
            torch.cuda.synchronize()
        if watchdog is not None:
            watchdog.cancel()

    def step_device():
        if graphed is not None:
            return graphed.replay()[0]
        return step_eager()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        last = None
        for _ in range(steps):
            last = fn()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = t.item()
        return ms, last

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()          # sampled across warm-up, the timed region, the profiled pass and the e2e region
    for _ in range(max(args.warmup, 3)):
        step_device()
    launches0 = lib.clica_launch_count(-1)
    ms_total, last = timed(step_device, args.steps)
    launches = lib.clica_launch_count(-1) - launches0
    if graphed is not None:
        launches = graphed.launches_per_replay * args.steps     # recorded once, executed once per replay
    ms_step = ms_total / args.steps
    value = B_global / (ms_step * 1e-3)
    loss_value = float(last.item())

    # ---- per-kernel-family device time over the same steps (CUDA events inside the library) ------------
    import ctypes
    prof_steps = min(args.steps, 20)
    _lib.check(lib.clica_prof_enable(1), "clica_prof_enable")
    for _ in range(prof_steps):
        step_eager()                                            # (events cannot bracket kernels inside a graph)
    ms_f = (ctypes.c_float * 7)()
    n_f = (ctypes.c_int * 7)()
    _lib.check(lib.clica_prof_collect(ms_f, n_f), "clica_prof_collect")
    lib.clica_prof_enable(0)
    fam_ms = {name: ms_f[i] / prof_steps for i, name in enumerate(FAMILIES)}
    fam_n = {name: n_f[i] // prof_steps for i, name in enumerate(FAMILIES)}

    # ---- e2e: host buffers, H2D every step, loss read back every step, torch.optim.Adam as the script ---
    torch.manual_seed(0)
    f2 = encoders.get_mlp(n, n, [10 * n, 50 * n, 50 * n, 50 * n, 50 * n, 10 * n]).to(dev)
    opt2 = torch.optim.Adam(f2.parameters(), lr=1e-4)
    h2 = lambda z: f2(g(z))

    def step_e2e_eager():
        # exactly what the unchanged main_mlp.py does per step on the drop-in modules
        z1 = z1_h.to(dev, non_blocking=True)
        z2 = z2_h.to(dev, non_blocking=True)
        if world == 1:
            opt2.zero_grad()
            a, b = h2(z1), h2(z2)
            total, _, parts = crit(z1, z2, torch.roll(z1, 1, 0), a, b, torch.roll(a, 1, 0))
            total.backward()
            opt2.step()
        else:
            total, parts = sharded.sharded_train_step(f2, g, opt2, z1, z2, p, tau, 0.5)
        return total.item(), [float(x) for x in (parts.tolist() if torch.is_tensor(parts) else [q.item() for q in parts])]

    def time_e2e(fn):
        for _ in range(max(args.warmup, 3)):
            fn()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            fn()
        barrier()
        ms = (time.perf_counter() - t0) * 1e3 / args.steps
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = t.item()
        return ms

    e2e_eager_ms = time_e2e(step_e2e_eager)
    e2e_ms, e2e_mode = e2e_eager_ms, "eager drop-in modules (as main_mlp.py runs them)"
    if args.graph and (world == 1 or args.graph_multi):
        # public API for a host-fed loop: GraphedTrainStep(host_io=True).step_host(z1_host, z2_host) stages the
        # host latents in pinned memory; the graph copies them to the device, runs the step and copies
        # (loss, pos_mean, neg_mean) back; step_host waits for that copy and returns Python floats
        from clica_b200.graphed import GraphedTrainStep
        torch.manual_seed(0)
        f3 = encoders.get_mlp(n, n, [10 * n, 50 * n, 50 * n, 50 * n, 50 * n, 10 * n]).to(dev)
        graphed_io = GraphedTrainStep(f3, g, crit, B_local, n, lr=1e-4, host_io=True,
                                      group=(dist.group.WORLD if world > 1 else None))
        e2e_ms = time_e2e(lambda: graphed_io.step_host(z1_h, z2_h))
        e2e_mode = "GraphedTrainStep.step_host (CUDA graph incl. H2D of the batch and D2H of the loss scalars)"
    e2e_value = B_global / (e2e_ms * 1e-3)
    clocks = sampler.stop() if rank == 0 else None

    if rank == 0:
        peaks = load_peaks()
        work = algorithmic_work(wl, B_local, B_global, world)
        sm_count = torch.cuda.get_device_properties(dev).multi_processor_count
        f_sm = (clocks["sm_mhz"] or 1965.0) * 1e6
        gemm_ms = fam_ms["gemm_tc"] + fam_ms["gemm_simt"]
        loss_ms = fam_ms["loss_fwd"] + fam_ms["loss_bwd"]
        kernels = {
            "encoder_gemm": {"ms_per_step": gemm_ms, "launches_per_step": fam_n["gemm_tc"] + fam_n["gemm_simt"],
                             "tflops": work["enc_flops"] / (gemm_ms * 1e-3) / 1e12 if gemm_ms > 0 else None,
                             "tc_ms": fam_ms["gemm_tc"], "simt_ms": fam_ms["gemm_simt"]},
            "loss_fwd": {"ms_per_step": fam_ms["loss_fwd"],
                         "fp32_pipe_frac": work["loss_ops_fwd"] / (fam_ms["loss_fwd"] * 1e-3 * sm_count * 128 * f_sm) if fam_ms["loss_fwd"] > 0 else None},
            "loss_bwd": {"ms_per_step": fam_ms["loss_bwd"],
                         "fp32_pipe_frac": work["loss_ops_bwd"] / (fam_ms["loss_bwd"] * 1e-3 * sm_count * 128 * f_sm) if fam_ms["loss_bwd"] > 0 else None},
            "loss_fused_hbm": {"bytes_min": work["loss_bytes"],
                               "gbs": work["loss_bytes"] / (loss_ms * 1e-3) / 1e9 if loss_ms > 0 else None,
                               "frac_of_hbm_peak": work["loss_bytes"] / (loss_ms * 1e-3) / 1e9 / peaks["hbm_gbs"] if loss_ms > 0 else None},
            "adam_ms": fam_ms["adam"], "loss_aux_ms": fam_ms["loss_aux"], "misc_ms": fam_ms["misc"],
            "sum_ms": sum(fam_ms.values()), "fp32_pipe_clock_mhz": f_sm / 1e6,
        }
        if gemm_ms >= loss_ms:
            ach = work["enc_flops"] / (gemm_ms * 1e-3) / 1e12 if gemm_ms > 0 else 0.0
            mma_per_product = 3 if args.gemm_mode == "3xtf32" else 1
            traffic, traffic_src = None, None
            tpath = os.path.join(ROOT, "profiles", "r1_traffic.json")     # from the committed ncu --set full capture
            if os.path.exists(tpath):
                with open(tpath) as fh:
                    tj = json.load(fh)
                if tj.get("workload") == args.workload and world == 1 and args.gemm_mode == "3xtf32":
                    traffic, traffic_src = tj["dram_bytes_per_launch"], tj["source"]
            roofline = {"bound": "tensor", "kernel": "encoder GEMMs (" + ("tcgen05 " + args.gemm_mode if fam_ms["gemm_tc"] > fam_ms["gemm_simt"] else "CUDA-core fp32") + ")",
                        "achieved": ach, "peak": peaks["bf16_tflops_sustained"], "unit": "TFLOP/s",
                        "frac": ach / peaks["bf16_tflops_sustained"], "traffic": traffic, "traffic_source": traffic_src,
                        "peak_source": peaks["source"] + " bf16 sustained; tf32 MMA peak is half of it and 3xtf32 issues 3 MMAs per product",
                        # the same achieved number in units of the instructions the mode actually issues: tf32 MMAs run
                        # at half the bf16 rate and the fp32-accurate 3xtf32 mode spends 3 of them per product
                        "tf32_mma_tflops_issued": ach * mma_per_product,
                        "frac_of_tf32_peak": ach * mma_per_product / (peaks["bf16_tflops_sustained"] / 2.0),
                        "launches_per_step": fam_n["gemm_tc"] + fam_n["gemm_simt"],
                        "avg_launch_us": gemm_ms * 1e3 / max(fam_n["gemm_tc"] + fam_n["gemm_simt"], 1),
                        "share_of_step": gemm_ms / max(kernels["sum_ms"], 1e-9)}
        else:
            ach = work["loss_bytes"] / (loss_ms * 1e-3) / 1e9
            roofline = {"bound": "hbm", "kernel": "fused Lp-InfoNCE fwd+bwd (compute-bound on the FP32 pipe; see kernels.loss_*.fp32_pipe_frac)",
                        "achieved": ach, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": ach / peaks["hbm_gbs"],
                        "traffic": None, "peak_source": peaks["source"],
                        "share_of_step": loss_ms / max(kernels["sum_ms"], 1e-9)}
        cpu, eager_gpu = None, None
        if world == 1 and not args.no_cpu_baseline:
            res = cpu_baseline_run(wl, steps=3, warmup=1, rows=args.cpu_sample_rows, budget_s=20.0)
            cpu = {k: res[k] for k in ("value", "unit", "cores", "kind", "sample")}
            eager_gpu = torch_eager_gpu_run(wl, dev)
        line = {
            "metric": "positive-pairs/sec", "value": value, "unit": "pairs/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": args.scaling,
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": wl["desc"], "n": n, "batch_per_gpu": B_local, "global_batch": B_global, "p": p,
                       "tau": tau, "gemm_mode": args.gemm_mode, "parallelism": f"row-sharded x{world}, all-gathered negatives" if world > 1 else "single GPU",
                       "l2_policy": "no L2 flush: the step rewrites >126 MB of activations/gradients per iteration (inputs larger than L2 at c2: 2 x 109 MB saved activations)"},
            "e2e": {"value": e2e_value, "unit": "pairs/s", "ms_per_step": e2e_ms,
                    "h2d_bytes_per_step": 2 * B_local * n * 4, "d2h_bytes_per_step": 12, "api": e2e_mode,
                    "eager_dropin_ms_per_step": e2e_eager_ms,
                    "eager_dropin_value": B_global / (e2e_eager_ms * 1e-3)},
            "step_mode": step_mode,
            "gpu_launches": int(launches), "launches_per_step": launches / args.steps,
            "clocks": clocks, "roofline": roofline, "kernels": kernels, "cpu_baseline": cpu,
            "torch_eager_gpu_baseline": eager_gpu, "loss": loss_value,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        if graphed is not None:
            # communicators whose collectives live in instantiated CUDA graphs do not tear down cleanly (the
            # process-group destructor waits on work the graphs still reference): every rank is done, leave
            barrier()
            sys.stdout.flush()
            sys.stderr.flush()
            os._exit(0)
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
