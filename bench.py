#!/usr/bin/env python
"""bench.py -- positive-pairs/sec of cl-ica's InfoNCE training step on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c2|c3]
                    [--scaling weak|strong] [--gemm-mode 3xtf32|tf32|fp32] [--no-c3]

A "step" is the body of main_mlp.py:258-285 (unsupervised branch): zero_grad -> h(z1) -> h(z2) -> roll ->
LpSimCLRLoss -> backward -> Adam.step, h = f o g with f the 7-layer MLP encoder and g a frozen 3-layer mixing
net, on synthetic latents of the named shape (c2: n=10, B=6144 per GPU, sphere, p=2, tau=1; c3: n=40,
B=8192, p=3).  One JSON line is printed by rank 0 (see DESIGN.md "Measurement" for every key).

  value     device-timed pairs/s with inputs resident in HBM (CUDA events, max over ranks)
  e2e       same step fed from pinned HOST buffers every step, loss scalars read back every step
  c3_strong (default workload only) BASELINE config 3 -- n=40, p=3, GLOBAL batch 8192 sharded over the N ranks with
            all-gathered negatives -- timed in the same process: the north star's 1/2/4/8-GPU scaling curve
  --impl reference: the reference's own CPU path on this box's host cores.  With baseline/_ref present (the read-only
            copy of the reference that `build()` places there) and a global batch the host can materialise, the step is
            the UNMODIFIED reference's losses.LpSimCLRLoss / encoders.get_mlp / spaces samplers at the full batch
            (kind "reference"); otherwise oracle/torch_port on a bounded sample of anchors against all negatives
            (kind "port").
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
L2_POLICY = ("no L2 flush: the step rewrites >126 MB of activations/gradients per iteration (inputs larger than L2 "
             "at c2: 2 x 109 MB saved activations)")


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
    ap.add_argument("--no-c3", action="store_true", help="skip the extra c3_strong measurement")
    ap.add_argument("--no-graph", dest="graph", action="store_false",
                    help="time the eager step instead of the CUDA-graph replay")
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


def load_fp32_pipe_peak():
    """Calibrated FP32-pipe peak (lane-op/s at the boost clock) from tools/fp32_pipe_probe.cu, when committed."""
    path = os.path.join(ROOT, "profiles", "r2_fp32_pipe_probe.json")
    if os.path.exists(path):
        try:
            with open(path) as fh:
                return json.load(fh)
        except Exception:
            return None
    return None


def config_of(wl, B_local, B_global):
    """Identical for both arms (--impl ours / reference): the workload and nothing implementation-specific."""
    return {"workload": wl["desc"], "n": wl["n"], "batch_per_gpu": B_local, "global_batch": B_global,
            "p": wl["p"], "tau": wl["tau"], "l2_policy": L2_POLICY}


# ---------------------------------------------------------------------------------------- CPU baselines ----
def cpu_port_run(wl, steps, warmup, rows=0, budget_s=25.0, M_global=None):
    """oracle.torch_port (the reference's torch formulation) on the host cores on a bounded sample: `rows` anchors /
    positives through encoder fwd+bwd, contrasted against ALL M_global negatives."""
    import torch
    from oracle import torch_port as tp
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    n, B, p, tau = wl["n"], wl["B"], wl["p"], wl["tau"]
    M = M_global or B
    torch.manual_seed(0)
    f = tp.build_encoder(n)
    g = tp.build_mixing(n, 3, seed=0)
    opt = torch.optim.Adam(f.parameters(), lr=1e-4)
    z1, z2 = tp.synth_latents(M, n, wl["space"], seed=0)
    with torch.no_grad():
        z3_rec = f(g(z1)).detach().clone().requires_grad_(True)      # all M negatives, pre-encoded
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
                sample=f"{rows} of {M} anchor/positive pairs per step (encoder fwd+bwd on the sample, loss "
                       f"fwd+bwd against all {M} negatives, Adam); {steps} steps, torch {torch.__version__} CPU",
                ms_per_step=dt * 1e3, rows=rows)


def cpu_reference_run(wl, steps, warmup, budget_s=150.0):
    """The UNMODIFIED reference (baseline/_ref) at the full batch on the host cores: its own spaces / latent_spaces
    samplers (outside the timed region, as SURVEY 8d prescribes), encoders.get_mlp, losses.LpSimCLRLoss,
    torch.optim.Adam, in the order of main_mlp.py:258-285.  Returns None when baseline/_ref is absent or the batch
    cannot be materialised on this host (the reference needs ~12 B^2 d bytes)."""
    import torch
    from clica_b200 import vendor, synth
    ref = vendor.vendored_dir()
    n, B, p, tau = wl["n"], wl["B"], wl["p"], wl["tau"]
    if ref is None or wl["space"] != "sphere" or 12.0 * B * B * n * 4 > 48e9:
        return None
    import importlib.util
    import warnings

    def load(name):
        spec = importlib.util.spec_from_file_location("_bench_ref_" + name, os.path.join(ref, name + ".py"))
        mod = importlib.util.module_from_spec(spec)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore", SyntaxWarning)
            spec.loader.exec_module(mod)
        return mod
    sys.path.insert(0, ref)          # the reference modules import each other by bare name
    try:
        losses, encoders, spaces, latent_spaces = load("losses"), load("encoders"), load("spaces"), load("latent_spaces")
    finally:
        sys.path.remove(ref)
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    torch.manual_seed(0)
    space = spaces.NSphereSpace(n)
    latent = latent_spaces.LatentSpace(space=space,
                                       sample_marginal=lambda space, size, device="cpu": space.uniform(size, device=device),
                                       sample_conditional=lambda space, z, size, device="cpu": space.normal(z, 0.05, size, device))
    z1 = latent.sample_marginal(size=B, device="cpu")
    z2 = latent.sample_conditional(z1, size=B, device="cpu")
    g = synth.build_mixing(n, 3, seed=0)        # frozen input transform; the reference's own search takes minutes
    f = encoders.get_mlp(n_in=n, n_out=n, layers=[n * 10, n * 50, n * 50, n * 50, n * 50, n * 10])
    crit = losses.LpSimCLRLoss(p=p, tau=tau, simclr_compatibility_mode=True)
    opt = torch.optim.Adam(f.parameters(), lr=1e-4)
    h = lambda z: f(g(z))

    def step():                                  # main_mlp.py:258-285, unsupervised branch
        z3 = torch.roll(z1, 1, 0)
        opt.zero_grad()
        z1_rec = h(z1)
        z2_rec = h(z2)
        z3_rec = torch.roll(z1_rec, 1, 0)
        total, _, parts = crit(z1, z2, z3, z1_rec, z2_rec, z3_rec)
        total.backward()
        opt.step()
        return total.item(), [x.item() for x in parts]

    t0 = time.perf_counter()
    step()
    first = time.perf_counter() - t0
    n_w = max(0, warmup - 1)
    if first * (n_w + steps) > budget_s:          # slow host: keep the whole run inside the budget
        n_w = 0
        steps = max(1, int(budget_s / first))
    for _ in range(n_w):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / steps
    return dict(value=B / dt, unit="pairs/s", cores=cores, kind="reference",
                sample=f"full batch: {B} of {B} pairs per step through the unmodified reference modules of baseline/_ref "
                       f"(losses.LpSimCLRLoss, encoders.get_mlp, spaces.NSphereSpace samplers, torch.optim.Adam; step order "
                       f"of main_mlp.py:258-285); {steps} timed steps, torch {torch.__version__} CPU",
                ms_per_step=dt * 1e3, rows=B, steps=steps)


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
    world = max(1, args.gpus)
    if args.scaling == "weak":
        B_local, B_global = wl["B"], wl["B"] * world
    else:
        B_local, B_global = wl["B"] // world, wl["B"]
    res = None
    if args.cpu_sample_rows <= 0 and B_global == wl["B"]:
        res = cpu_reference_run(wl, max(1, args.steps), max(0, args.warmup))
    if res is None:
        res = cpu_port_run(wl, max(1, args.steps), max(0, args.warmup), args.cpu_sample_rows, budget_s=120.0,
                           M_global=B_global)
    line = {
        "impl": "reference", "metric": "positive-pairs/sec", "value": res["value"], "unit": "pairs/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": res["ms_per_step"],
        "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config_of(wl, B_local, B_global),
        "parallelism": "host CPU threads (torch intra-op)",
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


class Ctx:
    pass


def family_of(kernel_name):
    n = kernel_name
    if "lpnce_fwd" in n:
        return "loss_fwd"
    if "lpnce_bwd" in n:
        return "loss_bwd"
    if "lpnce_" in n:
        return "loss_aux"
    if "gemm_tc_kernel" in n:
        return "gemm_tc"
    if "skinny" in n or "gemm_simt" in n or "mixing" in n:
        return "gemm_simt"
    if "adam" in n:
        return "adam"
    if "split_planes" in n or "colsum" in n or "sampler" in n:
        return "misc"
    if "nccl" in n.lower():
        return "nccl"
    if "memcpy" in n.lower() or "memset" in n.lower():
        return "memops"
    return "torch_other"


def profile_families(step_fn, k):
    """Per-family kernel time (ms per step) and launches per step from CUPTI kernel records of k calls of step_fn."""
    import torch
    from torch.profiler import profile, ProfilerActivity
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(k):
            step_fn()
        torch.cuda.synchronize()
    ms = {name: 0.0 for name in FAMILIES + ["nccl", "torch_other", "memops"]}
    cnt = {name: 0 for name in ms}
    seen = 0
    for ev in prof.events():
        dt = getattr(ev, "device_time_total", None)
        if dt is None:
            dt = getattr(ev, "cuda_time_total", 0.0)
        if str(getattr(ev, "device_type", "")).endswith("CUDA") and dt and dt > 0:
            fam = family_of(ev.name)
            ms[fam] += dt / 1e3
            cnt[fam] += 1
            seen += 1
    if seen == 0:
        raise RuntimeError("no CUDA kernel records")
    return {a: b / k for a, b in ms.items()}, {a: b // k for a, b in cnt.items()}


def measure_workload(cx, wl_name, scaling, steps, warmup, full):
    """Times one workload on the ranks of `cx`; returns a dict (rank 0 uses it).  `full` adds the eager drop-in e2e."""
    import torch
    import torch.distributed as dist
    from clica_b200 import _lib, sharded
    from clica_b200.optim import FusedAdam
    from clica_b200 import synth as tp
    from clica_b200.graphed import GraphedTrainStep
    import encoders
    import losses
    args, world, rank, dev, lib = cx.args, cx.world, cx.rank, cx.dev, cx.lib

    wl = WORKLOADS[wl_name]
    n, p, tau = wl["n"], wl["p"], wl["tau"]
    if scaling == "weak":
        B_local, B_global = wl["B"], wl["B"] * world
    else:
        assert wl["B"] % world == 0
        B_local, B_global = wl["B"] // world, wl["B"]

    torch.manual_seed(0)                                   # identical initial weights on every rank
    f = encoders.get_mlp(n, n, [10 * n, 50 * n, 50 * n, 50 * n, 50 * n, 10 * n]).to(dev)
    g = tp.build_mixing(n, 3, seed=0).to(dev)
    crit = losses.LpSimCLRLoss(p=p, tau=tau, simclr_compatibility_mode=True)
    z1_h, z2_h = tp.synth_latents(B_global, n, wl["space"], seed=0)
    z1_h = z1_h[rank * B_local:(rank + 1) * B_local].contiguous().pin_memory()
    z2_h = z2_h[rank * B_local:(rank + 1) * B_local].contiguous().pin_memory()
    z1_d, z2_d = z1_h.to(dev), z2_h.to(dev)
    h = lambda z: f(g(z))
    group = dist.group.WORLD if world > 1 else None

    # ---- the step, device-resident flavour (value): fused Adam, no host sync inside the loop ------------
    opt = FusedAdam(f.parameters(), lr=1e-4)
    # anchors and positives go through the encoder as ONE 2B-row batch (same arithmetic per row; the GEMMs get
    # twice the rows per launch); the eager drop-in e2e flavour below keeps the script's two separate calls
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

    # the same step recorded once into a CUDA graph (clica_b200.graphed.GraphedTrainStep) and replayed -- no host
    # work between the kernels of a step.  CLICA_GRAPH=0 times the eager step instead.
    step_mode = "eager"
    graphed = None
    use_graph = args.graph and (world == 1 or args.graph_multi)
    if use_graph:
        watchdog = None
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
            graphed = GraphedTrainStep(f, g, crit, B_local, n, lr=1e-4, host_io=False, group=group)
            graphed.stage(z1_d, z2_d)
            graphed.replay()
            torch.cuda.synchronize()
            step_mode = "cuda_graph"
            cx.any_multi_graph = cx.any_multi_graph or world > 1
        except Exception as exc:           # same kernels either way: time the eager step and say so in the JSON line
            if world > 1:
                raise                      # ranks must not diverge on the step flavour
            sys.stderr.write(f"bench.py: CUDA-graph step unavailable ({exc!r}); timing the eager step\n")
            graphed = None
            use_graph = False
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

    def timed(fn, k):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        last = None
        for _ in range(k):
            last = fn()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = t.item()
        return ms, last

    for _ in range(max(warmup, 3)):
        step_device()
    launches0 = lib.clica_launch_count(-1)
    ms_total, last = timed(step_device, steps)
    launches = lib.clica_launch_count(-1) - launches0
    if graphed is not None:
        launches = graphed.launches_per_replay * steps     # recorded once, executed once per replay
    ms_step = ms_total / steps
    value = B_global / (ms_step * 1e-3)
    loss_value = float(last.item())

    # ---- per-kernel-family device time ---------------------------------------------------------------
    # Preferred: CUPTI kernel records (torch.profiler) of the SAME graph replays that were timed -- exact per-kernel
    # durations inside the timed step flavour, torch's own kernels and NCCL included.  Fallback (profiler unavailable,
    # e.g. when the whole process runs under ncu): the library's CUDA events around its own launches in an eager pass
    # queued behind a device-side spin (kernel time without host launch gaps; ~3 us of event overhead per launch).
    fam_ms, fam_n, fam_src = None, None, None
    # With programmatic dependent launch (CLICA_PDL, default on) a kernel's CTAs are resident -- and its CUPTI record
    # runs -- while the tail of its predecessor drains, so per-kernel durations of the timed graph overlap.  The
    # breakdown therefore comes from the SAME step recorded once more with plain stream order (identical kernels).
    prof_step, prof_note = step_device, ""
    pdl_on = os.environ.get("CLICA_PDL", "1") != "0"
    graphed_prof = None
    if graphed is not None and pdl_on:
        os.environ["CLICA_PDL"] = "0"
        try:
            torch.manual_seed(0)
            f_p = encoders.get_mlp(n, n, [10 * n, 50 * n, 50 * n, 50 * n, 50 * n, 10 * n]).to(dev)
            graphed_prof = GraphedTrainStep(f_p, g, crit, B_local, n, lr=1e-4, host_io=False, group=group)
            graphed_prof.stage(z1_d, z2_d)
            with torch.no_grad():              # same data regime as the timed graph: start from its current weights
                for q_dst, q_src in zip(f_p.parameters(), f.parameters()):
                    q_dst.copy_(q_src)
            torch.autograd.graph.increment_version(list(f_p.parameters()))
            for _ in range(3):
                graphed_prof.replay()
            prof_step = lambda: graphed_prof.replay()[0]
            prof_note = ", recorded without programmatic dependent launch (per-kernel durations do not overlap)"
        finally:
            os.environ["CLICA_PDL"] = "1"
    try:
        fam_ms, fam_n = profile_families(prof_step, min(steps, 10))
        fam_src = ("torch.profiler (CUPTI) kernel records of the timed step flavour (" + step_mode.split(" ")[0] + ")"
                   + prof_note)
    except Exception as exc:
        sys.stderr.write(f"bench.py: torch.profiler breakdown unavailable ({exc!r}); using library events\n")
    if fam_ms is None:
        import ctypes
        prof_steps = min(steps, 10)
        spin = int(4e-3 * 1.9e9)
        step_eager()
        _lib.check(lib.clica_prof_enable(1), "clica_prof_enable")
        for _ in range(prof_steps):
            torch.cuda._sleep(spin)
            step_eager()
            torch.cuda.synchronize()
        ms_f = (ctypes.c_float * 7)()
        n_f = (ctypes.c_int * 7)()
        _lib.check(lib.clica_prof_collect(ms_f, n_f), "clica_prof_collect")
        lib.clica_prof_enable(0)
        fam_ms = {name: ms_f[i] / prof_steps for i, name in enumerate(FAMILIES)}
        fam_n = {name: n_f[i] // prof_steps for i, name in enumerate(FAMILIES)}
        fam_ms.update(nccl=0.0, torch_other=0.0, memops=0.0)
        fam_n.update(nccl=0, torch_other=0, memops=0)
        fam_src = "library CUDA events around each launch, eager pass behind a device-side spin"

    # ---- e2e: host buffers, H2D every step, loss read back every step -----------------------------------
    def time_e2e(fn):
        for _ in range(max(warmup, 3)):
            fn()
        barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            fn()
        barrier()
        ms = (time.perf_counter() - t0) * 1e3 / steps
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = t.item()
        return ms

    e2e_eager_ms = None
    if full:
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
        e2e_eager_ms = time_e2e(step_e2e_eager)
        del f2, opt2

    if use_graph:
        # public API for a host-fed loop: GraphedTrainStep(host_io=True).step_host(z1_host, z2_host) stages the
        # host latents in pinned memory; the graph copies them to the device, runs the step and copies
        # (loss, pos_mean, neg_mean) back; step_host waits for that copy and returns Python floats
        torch.manual_seed(0)
        f3 = encoders.get_mlp(n, n, [10 * n, 50 * n, 50 * n, 50 * n, 50 * n, 10 * n]).to(dev)
        graphed_io = GraphedTrainStep(f3, g, crit, B_local, n, lr=1e-4, host_io=True, group=group)
        # (a) the producer of the batch writes it straight into the step's pinned input views (pinned_inputs()); every
        #     step = graph launch (H2D of the 2 x B x n batch from that pinned memory, the step, D2H of the three loss
        #     scalars) + stream sync + float conversion;  (b) the batch lives in the caller's own pinned tensors and
        #     step_host(z1, z2) first copies it host-to-host into the staging views
        e2e_staged_ms = time_e2e(lambda: graphed_io.step_host(z1_h, z2_h))
        pin1, pin2 = graphed_io.pinned_inputs()
        pin1.copy_(z1_h)
        pin2.copy_(z2_h)
        e2e_ms = time_e2e(lambda: graphed_io.step_host())
        e2e_mode = ("GraphedTrainStep.pinned_inputs() + step_host(): CUDA graph incl. the H2D copy of the batch from pinned host "
                    "memory and the D2H copy of the loss scalars; staged_ms_per_step = step_host(z1, z2) from the caller's own "
                    "pinned tensors (one more host-to-host copy)")
    else:
        if e2e_eager_ms is None:
            raise RuntimeError("no e2e flavour available")
        e2e_ms, e2e_mode = e2e_eager_ms, "eager drop-in modules (as main_mlp.py runs them)"

    work = algorithmic_work(wl, B_local, B_global, world)
    res = dict(wl=wl, B_local=B_local, B_global=B_global, ms_per_step=ms_step, value=value, loss=loss_value,
               step_mode=step_mode, launches=int(launches), launches_per_step=launches / steps,
               fam_ms=fam_ms, fam_n=fam_n, fam_src=fam_src, work=work,
               e2e={"value": B_global / (e2e_ms * 1e-3), "unit": "pairs/s", "ms_per_step": e2e_ms,
                    "h2d_bytes_per_step": 2 * B_local * n * 4, "d2h_bytes_per_step": 12, "api": e2e_mode})
    if use_graph:
        res["e2e"]["staged_ms_per_step"] = e2e_staged_ms
    if e2e_eager_ms is not None:
        res["e2e"]["eager_dropin_ms_per_step"] = e2e_eager_ms
        res["e2e"]["eager_dropin_value"] = B_global / (e2e_eager_ms * 1e-3)
    del graphed, graphed_prof, f, opt
    torch.cuda.empty_cache()
    return res


def kernel_report(res, sm_count, f_sm, peaks, fp32_probe):
    fam_ms, fam_n, work = res["fam_ms"], res["fam_n"], res["work"]
    gemm_ms = fam_ms["gemm_tc"] + fam_ms["gemm_simt"]
    loss_ms = fam_ms["loss_fwd"] + fam_ms["loss_bwd"]
    pipe_peak = sm_count * 128 * f_sm                       # nominal lane-op/s at the sampled clock
    if fp32_probe and fp32_probe.get("lane_ops_per_s"):
        pipe_peak = fp32_probe["lane_ops_per_s"] * (f_sm / (fp32_probe.get("sm_mhz", f_sm / 1e6) * 1e6))
    frac = lambda ops, ms: (ops / (ms * 1e-3 * pipe_peak)) if ms > 0 else None
    total = sum(fam_ms.values())
    return {
        "encoder_gemm": {"ms_per_step": gemm_ms, "launches_per_step": fam_n["gemm_tc"] + fam_n["gemm_simt"],
                         "tflops": work["enc_flops"] / (gemm_ms * 1e-3) / 1e12 if gemm_ms > 0 else None,
                         "tc_ms": fam_ms["gemm_tc"], "simt_ms": fam_ms["gemm_simt"]},
        "loss_fwd": {"ms_per_step": fam_ms["loss_fwd"], "fp32_pipe_frac": frac(work["loss_ops_fwd"], fam_ms["loss_fwd"])},
        "loss_bwd": {"ms_per_step": fam_ms["loss_bwd"], "fp32_pipe_frac": frac(work["loss_ops_bwd"], fam_ms["loss_bwd"])},
        "loss_fused_hbm": {"bytes_min": work["loss_bytes"],
                           "gbs": work["loss_bytes"] / (loss_ms * 1e-3) / 1e9 if loss_ms > 0 else None,
                           "frac_of_hbm_peak": work["loss_bytes"] / (loss_ms * 1e-3) / 1e9 / peaks["hbm_gbs"] if loss_ms > 0 else None},
        "adam_ms": fam_ms["adam"], "loss_aux_ms": fam_ms["loss_aux"], "misc_ms": fam_ms["misc"],
        "nccl_ms": fam_ms.get("nccl", 0.0), "torch_other_ms": fam_ms.get("torch_other", 0.0), "memops_ms": fam_ms.get("memops", 0.0),
        "sum_ms": total, "fp32_pipe_clock_mhz": f_sm / 1e6,
        "fp32_pipe_peak_source": ("profiles/r2_fp32_pipe_probe.json (FFMA microbenchmark), scaled to the sampled clock"
                                  if fp32_probe and fp32_probe.get("lane_ops_per_s") else "nominal SMs x 128 lanes x clock"),
        "timing": res.get("fam_src"),
    }


def run_ours(args):
    import torch
    import torch.distributed as dist
    import clica_b200
    from clica_b200 import _lib

    cx = Ctx()
    cx.args = args
    cx.world = int(os.environ.get("WORLD_SIZE", "1"))
    cx.rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if cx.world != args.gpus:
        if cx.world == 1 and args.gpus > 1:
            raise SystemExit("launch multi-GPU runs with torch.distributed.run (see the module docstring)")
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback for the product path)"
    torch.cuda.set_device(local_rank)
    cx.dev = torch.device("cuda", local_rank)
    cx.lib = _lib.load()
    cx.any_multi_graph = False
    os.environ["CLICA_GEMM_MODE"] = args.gemm_mode
    world, rank, dev = cx.world, cx.rank, cx.dev
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    sys.path.insert(0, clica_b200.DROPIN_DIR)

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()          # sampled across warm-up, the timed regions, the profiled passes and the e2e regions
    main = measure_workload(cx, args.workload, args.scaling, args.steps, args.warmup, full=True)
    c3 = None
    if args.workload == "c2" and args.scaling == "weak" and not args.no_c3 and WORKLOADS["c3"]["B"] % world == 0:
        c3 = measure_workload(cx, "c3", "strong", args.steps, args.warmup, full=False)
    clocks = sampler.stop() if rank == 0 else None

    if rank == 0:
        peaks = load_peaks()
        fp32_probe = load_fp32_pipe_peak()
        sm_count = torch.cuda.get_device_properties(dev).multi_processor_count
        f_sm = (clocks["sm_mhz"] or 1965.0) * 1e6
        wl = main["wl"]
        kernels = kernel_report(main, sm_count, f_sm, peaks, fp32_probe)
        fam_ms, fam_n, work = main["fam_ms"], main["fam_n"], main["work"]
        gemm_ms = fam_ms["gemm_tc"] + fam_ms["gemm_simt"]
        loss_ms = fam_ms["loss_fwd"] + fam_ms["loss_bwd"]
        if gemm_ms >= loss_ms:
            ach = work["enc_flops"] / (gemm_ms * 1e-3) / 1e12 if gemm_ms > 0 else 0.0
            mma_per_product = 3 if args.gemm_mode == "3xtf32" else 1
            traffic, traffic_src = None, None
            for name in ("r2_traffic.json", "r1_traffic.json"):          # from the committed ncu --set full capture
                tpath = os.path.join(ROOT, "profiles", name)
                if os.path.exists(tpath):
                    with open(tpath) as fh:
                        tj = json.load(fh)
                    if tj.get("workload") == args.workload and world == 1 and args.gemm_mode == "3xtf32":
                        traffic, traffic_src = tj["dram_bytes_per_launch"], tj["source"]
                    break
            n_l = fam_n["gemm_tc"] + fam_n["gemm_simt"]
            roofline = {"bound": "tensor", "kernel": "encoder GEMMs (" + ("tcgen05 " + args.gemm_mode if fam_ms["gemm_tc"] > fam_ms["gemm_simt"] else "CUDA-core fp32") + ")",
                        "achieved": ach, "peak": peaks["bf16_tflops_sustained"], "unit": "TFLOP/s",
                        "frac": ach / peaks["bf16_tflops_sustained"], "traffic": traffic, "traffic_source": traffic_src,
                        "peak_source": peaks["source"] + " bf16 sustained; tf32 MMA peak is half of it and 3xtf32 issues 3 MMAs per product",
                        # the same achieved number in units of the instructions the mode actually issues: tf32 MMAs run
                        # at half the bf16 rate and the fp32-accurate 3xtf32 mode spends 3 of them per product
                        "tf32_mma_tflops_issued": ach * mma_per_product,
                        "frac_of_tf32_peak": ach * mma_per_product / (peaks["bf16_tflops_sustained"] / 2.0),
                        "launches_per_step": n_l, "avg_launch_us": gemm_ms * 1e3 / max(n_l, 1),
                        "share_of_step": gemm_ms / max(main["ms_per_step"], 1e-9)}
        else:
            ach = work["loss_bytes"] / (loss_ms * 1e-3) / 1e9
            roofline = {"bound": "hbm", "kernel": "fused Lp-InfoNCE fwd+bwd (compute-bound on the FP32 pipe; see kernels.loss_*.fp32_pipe_frac)",
                        "achieved": ach, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": ach / peaks["hbm_gbs"],
                        "traffic": None, "peak_source": peaks["source"],
                        "share_of_step": loss_ms / max(main["ms_per_step"], 1e-9)}
        cpu, eager_gpu = None, None
        if world == 1 and not args.no_cpu_baseline:
            res = None
            if args.cpu_sample_rows <= 0:
                res = cpu_reference_run(wl, steps=3, warmup=1, budget_s=25.0)
            if res is None:
                res = cpu_port_run(wl, steps=3, warmup=1, rows=args.cpu_sample_rows, budget_s=20.0)
            cpu = {k: res[k] for k in ("value", "unit", "cores", "kind", "sample")}
            eager_gpu = torch_eager_gpu_run(wl, dev)
        line = {
            "metric": "positive-pairs/sec", "value": main["value"], "unit": "pairs/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": main["ms_per_step"], "higher_is_better": True, "scaling": args.scaling,
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config_of(wl, main["B_local"], main["B_global"]),
            "gemm_mode": args.gemm_mode,
            "parallelism": f"row-sharded x{world}, all-gathered negatives" if world > 1 else "single GPU",
            "e2e": main["e2e"], "step_mode": main["step_mode"],
            "gpu_launches": main["launches"], "launches_per_step": main["launches_per_step"],
            "clocks": clocks, "roofline": roofline, "kernels": kernels, "cpu_baseline": cpu,
            "torch_eager_gpu_baseline": eager_gpu, "loss": main["loss"],
        }
        if c3 is not None:
            # BASELINE config 3 = the north star's scaling config: GLOBAL batch 8192 at every N (strong scaling)
            line["c3_strong"] = {
                "config": config_of(c3["wl"], c3["B_local"], c3["B_global"]), "scaling": "strong",
                "ms_per_step": c3["ms_per_step"], "value": c3["value"], "unit": "pairs/s", "e2e": c3["e2e"],
                "step_mode": c3["step_mode"], "launches_per_step": c3["launches_per_step"], "loss": c3["loss"],
                "kernels": kernel_report(c3, sm_count, f_sm, peaks, fp32_probe),
            }
        print(json.dumps(line), flush=True)
    if world > 1:
        if cx.any_multi_graph:
            # communicators whose collectives live in instantiated CUDA graphs do not tear down cleanly (the
            # process-group destructor waits on work the graphs still reference): every rank is done, leave
            dist.barrier()
            torch.cuda.synchronize()
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
