"""Training-step benchmark (BASELINE config 5 shape, on the model train.py trains): hicedrn_Diff(self_condition=True),
GaussianDiffusion(loss_type='l2', 'linear'), B tiles per GPU, torch Adam(lr=2e-5) -- the loop of train.py:109-136 verbatim.
One JSON line: tiles/s over all ranks, ms per step (device time, max over ranks), per-family profile of rank 0.
  python scripts/bench_train.py --batch 64 --steps 10            (N > 1: launch under torch.distributed.run)"""
import argparse
import json
import os
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

import torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--blocks", type=int, default=32)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--profile", action="store_true")
    ap.add_argument("--model", choices=["hicedrn", "unet"], default="hicedrn")
    ap.add_argument("--optim", choices=["torch", "fused"], default="torch", help="torch.optim.Adam (train.py verbatim) or hicdiff_b200.optim.Adam")
    ap.add_argument("--allreduce", choices=["overlap", "flat"], default="flat",
                    help="N > 1: bucketed all-reduce overlapped with the backward (Unet), or one all-reduce of the flat buffer after the step")
    a = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl")
    from hicdiff_b200 import train as T
    from hicdiff_b200.hicdiff_condition import GaussianDiffusion
    from hicdiff_b200.model.hicedrn_Diff import hicedrn_Diff
    from hicdiff_b200.synthetic import synthetic_tiles

    torch.manual_seed(0)
    if a.model == "unet":       # BASELINE config 5: conditional Unet p_losses
        from hicdiff_b200.hicdiff_condition import Unet

        net = Unet(dim=64, dim_mults=(1, 2, 4, 8), self_condition=True)
        diff = GaussianDiffusion(net, image_size=64, timesteps=1000, loss_type="l2", beta_schedule="sigmoid").cuda()
    else:
        net = hicedrn_Diff(number_resnet=a.blocks, self_condition=True)
        diff = GaussianDiffusion(net, image_size=64, timesteps=1000, loss_type="l2", beta_schedule="linear", auto_normalize=False).cuda()
    diff.train()
    if world > 1:
        T.enable_gradient_allreduce(net, overlap=a.allreduce == "overlap")
    optname = "fused Adam (hd_adam_step)" if a.optim == "fused" else "torch Adam"
    if a.optim == "fused":
        from hicdiff_b200.optim import Adam as FusedAdam

        opt = FusedAdam(diff.parameters(), lr=2e-5)
    else:
        opt = torch.optim.Adam(diff.parameters(), lr=2e-5)
    clean, noisy = synthetic_tiles(a.batch, seed=1234 + rank)
    x = [noisy.cuda(), clean.cuda()]
    torch.manual_seed(1 + rank)

    def step():
        loss = diff(x)
        loss.backward()
        opt.step()
        opt.zero_grad()
        return loss

    for _ in range(a.warmup):
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.time()
    e0.record()
    for _ in range(a.steps):
        loss = step()
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / a.steps], device="cuda")
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms)
    if rank == 0:
        fwd_flops = 14.564e9 if a.model == "unet" else 314.162e9 * a.blocks / 32.0   # per tile
        out = {"metric": "train_tiles_per_sec", "value": a.batch * world / (ms * 1e-3), "unit": "tiles/s", "n_gpus": world,
               "ms_per_step": ms, "steps": a.steps, "warmup": a.warmup, "wall_s": time.time() - t0, "loss": float(loss.detach()),
               "dtype": "bf16 activations / fp32 parameters and gradients", "data": "synthetic",
               "config": {"workload": (f"conditional Unet (dim 64, mults 1/2/4/8) p_losses l2 + backward + {optname}, batch {a.batch}/GPU" if a.model == "unet" else
                                       f"hicedrn_Diff({a.blocks} blocks, self_condition) p_losses l2 + backward + {optname}, batch {a.batch}/GPU"),
                          "allreduce": ("none" if world == 1 else
                                        "4 buckets in backward-completion order, NCCL on a communication stream overlapped with the backward"
                                        if (a.allreduce == "overlap" and a.model == "unet") else
                                        "one NCCL all-reduce of the flat fp32 gradient buffer after the step")},
               "model_tflops": 3 * fwd_flops * a.batch / (ms * 1e-3) / 1e12,
               "device_bytes": net._trainer.device_bytes(), "launch_groups": net._trainer.num_launch_groups()}
        if a.profile:
            prof = net._trainer.profile(3)
            top = prof.pop("_top", {})
            out["slowest_ops_ms"] = {k: round(v["ms"], 4) for k, v in top.items()}
            out["families"] = {k: {"ms": round(v["ms"], 4), "ops": v["ops"], "tflops": round(v["flops"] / max(v["ms"], 1e-9) / 1e9, 1)}
                               for k, v in prof.items()}
            out["families_total_ms"] = round(sum(v["ms"] for v in prof.values()), 3)
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
