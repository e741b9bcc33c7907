"""torchrun check: the NCCL row-sharded step equals the single-GPU full-batch step (loss + parameter gradients).

    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/sharded_check.py
"""
import os, sys
import torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import clica_b200
from clica_b200 import sharded, synth
sys.path.insert(0, clica_b200.DROPIN_DIR)
import encoders, losses

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
for (n, B, p) in [(10, 2048, 2), (5, 1024, 1), (40, 1024, 3)]:
    torch.manual_seed(0)
    f = encoders.get_mlp(n, n, [10 * n, 50 * n, 50 * n, 50 * n, 50 * n, 10 * n]).to(dev)
    g = synth.build_mixing(n, 3, seed=0).to(dev)
    z1, z2 = synth.synth_latents(B, n, "real", c_param=0.3, seed=1)
    z1, z2 = (z1 * 2).to(dev), (z2 * 2).to(dev)
    Bl = B // world
    sl = slice(rank * Bl, (rank + 1) * Bl)
    a, b = f(g(z1[sl])), f(g(z2[sl]))
    loss, _, parts = sharded.sharded_lp_infonce(a, b, p, 1.0, 0.5, True)
    loss.backward(retain_graph=True)
    sharded.allreduce_grads(list(f.parameters()))
    g_sh = [prm.grad.clone() for prm in f.parameters()]
    f.zero_grad()
    # same backward with the bucketed all-reduce issued from inside the encoder backward (overlapped with its GEMMs)
    with sharded.overlapped_grad_allreduce(f.parameters()):
        loss.backward()
    g_ov = [prm.grad.clone() for prm in f.parameters()]
    f.zero_grad()
    gm = max(t.abs().max().item() for t in g_sh)
    err_ov = max((x - y).abs().max().item() for x, y in zip(g_sh, g_ov)) / gm
    assert err_ov <= 1e-6, err_ov
    crit = losses.LpSimCLRLoss(p=p, tau=1.0, simclr_compatibility_mode=True)
    a2, b2 = f(g(z1)), f(g(z2))
    tot, _, parts2 = crit(None, None, None, a2, b2, torch.roll(a2, 1, 0))
    tot.backward()
    gmax = max(prm.grad.abs().max().item() for prm in f.parameters())
    err = max((gs - prm.grad).abs().max().item() for gs, prm in zip(g_sh, f.parameters())) / gmax
    if rank == 0:
        print(f"n={n} B={B} p={p} world={world}: loss sharded {loss.item():.7f} single {tot.item():.7f}  max grad err / max grad = {err:.2e}  (overlapped vs flat all-reduce: {err_ov:.1e})", flush=True)
    assert abs(loss.item() - tot.item()) <= 5e-6 * max(1.0, abs(tot.item())) and err <= 5e-4, (loss.item(), tot.item(), err)
# the CUDA-graph sharded step (NCCL collectives recorded into the graph) follows the eager sharded trajectory
if os.environ.get("CLICA_CHECK_GRAPH", "1") != "0":
    import copy
    from clica_b200.graphed import GraphedTrainStep
    from clica_b200.optim import FusedAdam
    n, B, p, lrate = 10, 2048, 2, 1e-3
    torch.manual_seed(0)
    f_g = encoders.get_mlp(n, n, [10 * n, 50 * n, 50 * n, 50 * n, 50 * n, 10 * n]).to(dev)
    f_e = copy.deepcopy(f_g)
    g = synth.build_mixing(n, 3, seed=0).to(dev)
    crit = losses.LpSimCLRLoss(p=p, tau=1.0, simclr_compatibility_mode=True)
    Bl = B // world
    step = GraphedTrainStep(f_g, g, crit, Bl, n, lr=lrate, host_io=False, group=dist.group.WORLD)
    opt = FusedAdam(f_e.parameters(), lr=lrate)
    for it in range(3):
        z1, z2 = synth.synth_latents(B, n, "sphere", seed=20 + it)
        sl = slice(rank * Bl, (rank + 1) * Bl)
        z1l, z2l = z1[sl].to(dev), z2[sl].to(dev)
        out = step(z1l, z2l).clone()
        le, parts = sharded.sharded_train_step(f_e, g, opt, z1l, z2l, p, 1.0, 0.5)
        if rank == 0:
            print(f"graphed sharded step {it}: loss {out[0].item():.7f} eager {le.item():.7f}", flush=True)
        assert abs(out[0].item() - le.item()) <= 1e-4 * max(1.0, abs(le.item())), (out, le)
dist.barrier()
torch.cuda.synchronize()
if rank == 0:
    print("sharded_check OK", flush=True)
# communicators referenced by instantiated CUDA graphs do not tear down cleanly: every rank is done, leave
sys.stdout.flush()
os._exit(0)
