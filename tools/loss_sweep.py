"""BASELINE.json config 5: fused Lp-distance + InfoNCE kernel sweep, B x d x p, on one GPU.

    python tools/loss_sweep.py [--out gpurun_out/loss_sweep.json] [--quick]
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/loss_sweep.py   (row-sharded)

For every point: forward and forward+backward device time (CUDA events on the launching stream, after
warm-up, L2 flushed between repetitions), the algorithmic-HBM GB/s (`24*B*d + 8*B` bytes fwd+bwd)
and the FP32-pipe fraction (`c_p*B*M*d` lane-ops forward, `(c_p+2)*B*M*d` backward) at the SM clock sampled
under load.  Inputs: z1 = randn(B,d), z2 = z1 + 0.05 randn, z3 = roll(z1,1,0), tau = 1.
"""
import argparse
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def sm_clock_mhz():
    try:
        out = subprocess.check_output(["nvidia-smi", "--query-gpu=clocks.sm", "--format=csv,noheader,nounits", "-i", "0"])
        return float(out.decode().strip().splitlines()[0])
    except Exception:
        return None


def main_sharded(args):
    """torchrun: row-sharded sweep (config 5 "at 1/2/4/8 GPUs").  Rank r owns B/W anchor rows against all B gathered
    columns: all-gather of the rows, local forward, all-gather of the row statistics, merged local backward --
    timed on the device, max over ranks.  (Written after round 1's GPU budget was spent; not yet run.)"""
    import torch
    import torch.distributed as dist
    import clica_b200
    from clica_b200 import sharded
    rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dev = torch.device("cuda", lr)
    dist.init_process_group("nccl", device_id=dev)
    clica_b200._lib.load()
    sms = torch.cuda.get_device_properties(dev).multi_processor_count
    rows = []
    for B in ([4096, 16384] if args.quick else [4096, 8192, 16384, 32768]):
        for d in ([10, 40] if args.quick else [10, 40, 128]):
            for p in (1, 2, 3):
                Bl = B // world
                g = torch.Generator(device="cpu").manual_seed(0)
                z1 = torch.randn(B, d, generator=g)[rank * Bl:(rank + 1) * Bl].to(dev)
                z2 = z1 + 0.05 * torch.randn(Bl, d, device=dev)
                gscale = torch.ones((), device=dev)

                def step():
                    z_all = sharded._all_gather_rows(z1, None)
                    li, lse, pos, rowstat = sharded.local_forward(z1, z2, z_all, p, 1.0, 0.5, True)
                    rs_all = sharded._all_gather_rows(rowstat, None)
                    return sharded.local_backward(z1, z2, z_all, rs_all, pos, rank * Bl, p, 1.0, 0.5, True, gscale)

                for _ in range(3):
                    step()
                reps = 3 if B * B * d > 2e10 * world else 10
                dist.barrier(); torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(reps):
                    step()
                e1.record()
                dist.barrier(); torch.cuda.synchronize()
                t = torch.tensor([e0.elapsed_time(e1) / reps], device=dev)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                ms = t.item()
                cp = {1: 2, 2: 2, 3: 3}[p]
                pe_rank = Bl * B * d
                rows.append(dict(B=B, d=d, p=p, n_gpus=world, fwd_bwd_ms=ms,
                                 fwd_bwd_fp32_pipe_frac_per_gpu=(2 * cp + 2) * pe_rank / (ms * 1e-3 * sms * 128 * 1.965e9),
                                 note="per-rank block B/W x B incl. both all-gathers; pipe fraction at the nominal 1965 MHz"))
                if rank == 0:
                    print(json.dumps(rows[-1]), flush=True)
    if rank == 0:
        os.makedirs(os.path.dirname(args.out), exist_ok=True)
        json.dump(dict(n_gpus=world, rows=rows), open(args.out, "w"), indent=1)
    dist.barrier()
    dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "loss_sweep.json"))
    ap.add_argument("--quick", action="store_true")
    args = ap.parse_args()
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        return main_sharded(args)
    import torch
    import clica_b200
    from clica_b200 import functional as F
    clica_b200._lib.load()
    dev = torch.device("cuda:0")
    sms = torch.cuda.get_device_properties(dev).multi_processor_count
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650.0}
    flush = torch.empty(192 << 20, dtype=torch.uint8, device=dev)       # > 126 MB L2
    Bs = [1024, 4096] if args.quick else [1024, 4096, 8192, 16384, 32768]
    ds = [10, 40] if args.quick else [10, 40, 128]
    rows = []
    for B in Bs:
        for d in ds:
            for p in (1, 2, 3):
                g = torch.Generator(device="cpu").manual_seed(0)
                z1 = torch.randn(B, d, generator=g).to(dev).requires_grad_(True)
                z2 = (z1.detach() + 0.05 * torch.randn(B, d, generator=g).to(dev)).requires_grad_(True)
                pe = B * B * d
                reps = 3 if pe > 2e10 else 10

                def fwd():
                    return F.lp_infonce(z1, z2, torch.roll(z1, 1, 0), p, 1.0, 0.5, True)[0]

                for _ in range(3):
                    fwd().backward()
                    z1.grad = z2.grad = None
                t_f = t_fb = 0.0
                clk = []
                for _ in range(reps):
                    flush.zero_()
                    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
                    e[0].record()
                    with torch.no_grad():
                        fwd()
                    e[1].record()
                    torch.cuda.synchronize()
                    flush.zero_()
                    e2 = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
                    e2[0].record()
                    fwd().backward()
                    e2[1].record()
                    c = sm_clock_mhz() if pe > 1e9 else None
                    torch.cuda.synchronize()
                    if c:
                        clk.append(c)
                    z1.grad = z2.grad = None
                    t_f += e[0].elapsed_time(e[1]) / reps
                    t_fb += e2[0].elapsed_time(e2[1]) / reps
                cp = {1: 2, 2: 2, 3: 3}[p]
                f_sm = (sorted(clk)[len(clk) // 2] if clk else 1965.0) * 1e6
                bytes_f, bytes_fb = 12 * B * d + 8 * B, 24 * B * d + 8 * B
                rows.append(dict(B=B, d=d, p=p, fwd_ms=t_f, fwd_bwd_ms=t_fb,
                                 fwd_hbm_gbs=bytes_f / (t_f * 1e-3) / 1e9, fwd_bwd_hbm_gbs=bytes_fb / (t_fb * 1e-3) / 1e9,
                                 fwd_bwd_hbm_frac=bytes_fb / (t_fb * 1e-3) / 1e9 / peaks["hbm_gbs"],
                                 fwd_fp32_pipe_frac=cp * pe / (t_f * 1e-3 * sms * 128 * f_sm),
                                 fwd_bwd_fp32_pipe_frac=(2 * cp + 2) * pe / (t_fb * 1e-3 * sms * 128 * f_sm),
                                 sm_mhz=f_sm / 1e6, reps=reps,
                                 note="times include torch.roll + autograd dispatch (launch-bound below ~0.05 ms)"))
                print(json.dumps(rows[-1]), flush=True)
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    json.dump(dict(gpu=torch.cuda.get_device_name(0), sms=sms, hbm_peak_gbs=peaks["hbm_gbs"], rows=rows), open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
