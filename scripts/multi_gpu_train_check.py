"""torchrun --nproc-per-node N scripts/multi_gpu_train_check.py
Data-parallel training (BASELINE config 5) on N GPUs: three steps of the conditional Unet with the BUCKETED, OVERLAPPED gradient
all-reduce and three with one flat all-reduce after the step, from the same seeds.  The first step's averaged gradients must agree
-- bit for bit at N = 2 (a sum of two numbers has one order), to fp32 summation-order level beyond that (NCCL picks its algorithm and
chunking by message size, so four buckets and one flat buffer add the eight contributions in different orders) -- and all ranks must
hold rank 0's parameters after the three steps."""
import os
import sys
from pathlib import Path

import torch
import torch.distributed as dist

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from hicdiff_b200 import train as T  # noqa: E402
from hicdiff_b200.hicdiff_condition import GaussianDiffusion, Unet  # noqa: E402
from hicdiff_b200.optim import Adam as FusedAdam  # noqa: E402
from hicdiff_b200.synthetic import synthetic_tiles  # noqa: E402


def run(overlap, rank, dev, B=8, steps=3):
    torch.manual_seed(0)
    net = Unet(dim=64, dim_mults=(1, 2, 4, 8), self_condition=True)
    diff = GaussianDiffusion(net, image_size=64, timesteps=1000, loss_type="l2", beta_schedule="sigmoid").to(dev)
    diff.train()
    T.enable_gradient_allreduce(net, overlap=overlap)
    opt = FusedAdam(diff.parameters(), lr=1e-3)
    clean, noisy = synthetic_tiles(B, seed=100 + rank)          # different data per rank: the all-reduce matters
    g = torch.Generator().manual_seed(7 + rank)
    losses = []
    first_grads = None
    for _ in range(steps):
        t = torch.randint(0, 1000, (B,), generator=g).to(dev)
        nz = torch.randn(B, 1, 64, 64, generator=g).to(dev)
        loss = diff.p_losses([noisy.to(dev), clean.to(dev)], t=t, noise=nz)
        loss.backward()
        if first_grads is None:
            first_grads = {k: p.grad.detach().clone() for k, p in net.named_parameters()}
        opt.step()
        opt.zero_grad()
        losses.append(float(loss.detach()))
    return {k: p.detach().clone() for k, p in net.named_parameters()}, losses, first_grads


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    pa, la, ga = run(True, rank, dev)
    pb, lb, gb = run(False, rank, dev)
    bitwise = all(torch.equal(ga[k], gb[k]) for k in ga)
    worst = max(float((ga[k] - gb[k]).abs().max() / gb[k].abs().max().clamp_min(1e-30)) for k in ga)   # relative to the tensor's scale
    same = bitwise if world <= 2 else worst <= 1e-5
    # replicas stay in sync: every rank holds rank 0's parameters
    sync = True
    for k in sorted(pa):
        ref = pa[k].clone()
        dist.broadcast(ref, src=0)
        sync = sync and torch.equal(ref, pa[k])
    flag = torch.tensor([1 if (same and sync) else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(f"multi_gpu_train_check world={world}: overlapped vs flat all-reduce, first-step gradients: ok {same} (bit-identical {bitwise}, "
              f"max |diff| / max |g| per tensor {worst:.1e}); "
              f"replicas in sync: {sync}; all ranks ok: {bool(flag.item())}; losses {['%.5f' % v for v in la]}", flush=True)
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
