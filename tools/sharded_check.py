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
    loss.backward()
    sharded.allreduce_grads(list(f.parameters()))
    g_sh = [prm.grad.clone() for prm in f.parameters()]
    f.zero_grad()
    crit = losses.LpSimCLRLoss(p=p, tau=1.0, simclr_compatibility_mode=True)
    a2, b2 = f(g(z1)), f(g(z2))
    tot, _, parts2 = crit(None, None, None, a2, b2, torch.roll(a2, 1, 0))
    tot.backward()
    gmax = max(prm.grad.abs().max().item() for prm in f.parameters())
    err = max((gs - prm.grad).abs().max().item() for gs, prm in zip(g_sh, f.parameters())) / gmax
    if rank == 0:
        print(f"n={n} B={B} p={p} world={world}: loss sharded {loss.item():.7f} single {tot.item():.7f}  max grad err / max grad = {err:.2e}", flush=True)
    assert abs(loss.item() - tot.item()) <= 5e-6 * max(1.0, abs(tot.item())) and err <= 5e-4, (loss.item(), tot.item(), err)
dist.barrier()
dist.destroy_process_group()
if rank == 0:
    print("sharded_check OK")
