"""SyncBatchNorm of the training kernels on N GPUs against one BatchNorm over all ranks' rows (torchrun, NCCL).
Run it under `timeout 120`: on the round-2 box the processes lingered after printing their result."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from pcdet_b200 import functional as F
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dev = torch.device("cuda", int(os.environ["LOCAL_RANK"])); torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev)
worst = 0.0
for dt in (torch.bfloat16, torch.float32):
    for c in (16, 64, 128):
        torch.manual_seed(7)
        sizes = [3000 + 777 * r for r in range(world)]                      # ragged: ranks hold different row counts
        y_all = (torch.randn(sum(sizes), c, device=dev) * 2 + 0.5).to(dt)
        go_all = torch.randn(sum(sizes), c, device=dev).to(dt)
        gamma, beta = torch.rand(c, device=dev) + 0.5, torch.randn(c, device=dev) * 0.2
        lo = sum(sizes[:rank]); hi = lo + sizes[rank]
        y, go = y_all[lo:hi].contiguous(), go_all[lo:hi].contiguous()
        rm, rv = torch.zeros(c, device=dev), torch.ones(c, device=dev)
        out, stats, sums = F.bn_train_fwd(y, gamma, beta, 1e-3, 0.01, rm, rv, relu=True, process_group=dist.group.WORLD)
        gy, gg, gb = F.bn_train_bwd(go, out, y, gamma, stats, relu=True, process_group=dist.group.WORLD, fwd_sums=sums)
        # one BatchNorm over every rank's rows (single-process kernels, tested against torch in tests/test_gpu_train_tc.py)
        rm2, rv2 = torch.zeros(c, device=dev), torch.ones(c, device=dev)
        o2, st2 = F.bn_train_fwd(y_all, gamma, beta, 1e-3, 0.01, rm2, rv2, relu=True)
        gy2, gg2, gb2 = F.bn_train_bwd(go_all, o2, y_all, gamma, st2, relu=True)
        gg_sum, gb_sum = gg.clone(), gb.clone()
        dist.all_reduce(gg_sum); dist.all_reduce(gb_sum)                    # parameter gradients are summed by DDP afterwards
        e = lambda a, b: float((a.float() - b.float()).abs().max() / b.float().abs().max().clamp_min(1e-9))
        errs = dict(out=e(out, o2[lo:hi]), gy=e(gy, gy2[lo:hi]), gg=e(gg_sum, gg2), gb=e(gb_sum, gb2), rm=e(rm, rm2), rv=e(rv, rv2))
        tol = {torch.bfloat16: 1e-2, torch.float32: 1e-5}[dt]          # bf16: one rounding flip of an output element is 3e-4
        worst = max(worst, *(v / tol for k, v in errs.items() if k in ("out", "gy")), *(v / 1e-5 for k, v in errs.items() if k not in ("out", "gy")))
        if rank == 0:
            print(f"syncbn {dt} c={c}: " + " ".join(f"{k}={v:.2e}" for k, v in errs.items()), flush=True)
t = torch.tensor([worst], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    print("WORST error / tolerance", float(t), "OK" if float(t) < 1.0 else "FAIL")
dist.destroy_process_group()
