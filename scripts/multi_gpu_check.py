"""torchrun --nproc-per-node N scripts/multi_gpu_check.py
Whole-genome pipeline (BASELINE config 4 shape, short chain) on N GPUs over NCCL: every rank denoises its shard of the
221 tiles, the finished tiles are all-gathered and reassembled; rank 0 then recomputes ALL tiles alone and checks that
the sharded result is bit-identical (Philox streams are keyed by the global tile id, so sharding must not matter)."""
import os
import sys
from pathlib import Path

import torch
import torch.distributed as dist

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from hicdiff_b200 import genome  # noqa: E402
from hicdiff_b200.hicdiff_condition import GaussianDiffusion  # noqa: E402
from hicdiff_b200.model.hicedrn_Diff import hicedrn_Diff  # noqa: E402
from hicdiff_b200.synthetic import synthetic_chromosome  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    sizes = (588, 632, 703, 802, 34, 588)
    mats = [synthetic_chromosome(n, seed=i).to(dev) for i, n in enumerate(sizes)]
    torch.manual_seed(0)
    net = hicedrn_Diff(self_condition=True)
    diff = GaussianDiffusion(net, image_size=64, timesteps=4, loss_type="l2", beta_schedule="sigmoid").to(dev)
    torch.manual_seed(123)
    out = genome.denoise_chromosomes(diff, mats, res=40000, max_batch=64)
    ok = True
    if rank == 0:
        # single-rank recomputation with the same seed: temporarily pretend there is no process group
        import hicdiff_b200.genome as G
        torch.manual_seed(123)
        seed = diff._next_seed(None)
        plan = diff._sync_plan()
        tiles = torch.cat([G.ops.tile_extract(m, 64, G.band_blocks_for(40000)) for m in mats])
        ref = torch.cat([plan.sample(min(64, tiles.shape[0] - b0), cond=tiles[b0:b0 + 64], seed=seed, tile_offset=b0)
                         for b0 in range(0, tiles.shape[0], 64)])
        k = 0
        for m, o in zip(mats, out):
            c = G.ops.tile_count(m.shape[0], 64, G.band_blocks_for(40000))
            r = G.ops.tile_scatter(ref[k:k + c], m.shape[0], 64, G.band_blocks_for(40000))
            ok = ok and torch.equal(o, r)
            k += c
        print(f"multi_gpu_check world={world}: sharded == single-rank: {ok}; tiles={tiles.shape[0]}", flush=True)
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.broadcast(flag, src=0)
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
