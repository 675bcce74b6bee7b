"""CPU tests of the host side: state_dict contract, loud failure without a GPU, C-ABI surface, sharding."""
import ctypes
import re
import subprocess
from pathlib import Path

import pytest
import torch

import helpers

ROOT = Path(__file__).resolve().parent.parent


def test_state_dict_layout_contract():
    """13 schedule buffers + the eps-net entries, names and shapes."""
    from hicdiff_b200.hicdiff_condition import GaussianDiffusion, Unet
    from hicdiff_b200.model.hicedrn_Diff import hicedrn_Diff

    d = GaussianDiffusion(Unet(dim=64, dim_mults=(1, 2, 4, 8)), image_size=64, timesteps=1000, loss_type="l2",
                          beta_schedule="sigmoid")
    sd = d.state_dict()
    assert len(sd) == 289
    assert all(sd[k].shape == (1000,) and sd[k].dtype == torch.float32 for k in list(sd)[:13])
    assert sd["model.init_conv.weight"].shape == (64, 2, 7, 7)
    assert sd["model.downs.0.3.1.weight"].shape == (64, 256, 1, 1)
    assert sd["model.downs.3.3.weight"].shape == (512, 256, 3, 3)
    assert sd["model.ups.0.3.1.weight"].shape == (256, 512, 3, 3)
    assert sd["model.mid_attn.fn.fn.to_out.weight"].shape == (512, 128, 1, 1)
    assert sd["model.downs.0.2.fn.fn.to_out.1.g"].shape == (1, 64, 1, 1)
    assert sd["model.final_res_block.mlp.1.weight"].shape == (128, 256)
    assert "model.downs.0.0.res_conv.weight" not in sd and "model.ups.0.0.res_conv.weight" in sd
    h = GaussianDiffusion(hicedrn_Diff(self_condition=True), image_size=64, timesteps=1000, loss_type="l2")
    hs = h.state_dict()
    assert len(hs) == 151
    assert hs["model.body.31.mlp.1.weight"].shape == (512, 1024) and hs["model.tail.weight"].shape == (1, 256, 3, 3)
    # strict round trip (what inference.py does with a checkpoint)
    d2 = GaussianDiffusion(Unet(dim=64, dim_mults=(1, 2, 4, 8)), image_size=64, timesteps=1000, loss_type="l2",
                           beta_schedule="sigmoid")
    d2.load_state_dict(sd, strict=True)
    # a T=1000 checkpoint does not load into a T=2000 object (buffers are in the state_dict)
    d3 = GaussianDiffusion(Unet(dim=64, dim_mults=(1, 2, 4, 8)), image_size=64, timesteps=2000, loss_type="l2",
                           beta_schedule="sigmoid")
    with pytest.raises(RuntimeError):
        d3.load_state_dict(sd, strict=True)


def test_sr3_variant_keys_and_attributes():
    from hicdiff_b200.hicdiff_sr3 import GaussianDiffusion, Unet

    net = Unet(dim=64, dim_mults=(1, 2, 4, 8), self_condition=True, noise_level_emb=True)
    sd = net.state_dict()
    assert "downs.0.0.noise_func.noise_func.0.weight" in sd and "downs.0.0.mlp.1.weight" not in sd
    assert sd["downs.0.0.noise_func.noise_func.0.weight"].shape == (64, 256)
    d = GaussianDiffusion(net, image_size=64, timesteps=1000)
    assert d.loss_type == "l2"                                   # SR3 defaults (hicdiff_sr3.py:499-501)
    assert "sqrt_alphas_cumprod_prev" not in d.state_dict()      # plain attribute, not a buffer
    assert d.sqrt_alphas_cumprod_prev.shape == (1001,) and d.sqrt_alphas_cumprod_prev.dtype == torch.float64
    for attr in ("channels", "out_dim", "self_condition", "random_or_learned_sinusoidal_cond"):
        assert hasattr(net, attr)


def test_reference_quirks_are_reproduced():
    from hicdiff_b200 import hicdiff, hicdiff_condition

    with pytest.raises(ValueError):
        hicdiff_condition.GaussianDiffusion(hicdiff_condition.Unet(dim=64), image_size=64, beta_schedule="nope")
    assert hicdiff.Unet(dim=64).self_condition is False           # default differs per file (hicdiff.py:263)
    assert hicdiff_condition.Unet(dim=64).self_condition is True
    with pytest.raises(NotImplementedError):                       # hicdiff.py crashes with self_condition=True (C.12)
        hicdiff.GaussianDiffusion(hicdiff.Unet(dim=64, self_condition=True), image_size=64)


def test_no_cpu_fallback():
    """The product path must fail loudly off-GPU: no ATen / oracle route exists behind the modules."""
    from hicdiff_b200.hicdiff_condition import GaussianDiffusion, Unet

    net = Unet(dim=64, dim_mults=(1, 2, 4, 8))
    x = torch.zeros(1, 1, 64, 64)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        net(x, torch.zeros(1, dtype=torch.long), x)
    d = GaussianDiffusion(net, image_size=64, timesteps=10, loss_type="l2")
    with pytest.raises(RuntimeError):
        d.super_resolution(x)
    with pytest.raises(RuntimeError, match="only holds parameters"):
        net.downs[0][0](x)
    # training: loss = diffusion(x) on a CPU module must raise too (the trainer is sm_100a-only), for both trainers
    from hicdiff_b200.model.hicedrn_Diff import hicedrn_Diff

    for m in (net, hicedrn_Diff(number_resnet=1, self_condition=True)):
        dd = GaussianDiffusion(m, image_size=64, timesteps=10, loss_type="l2")
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            dd([x, x])


def test_product_package_never_imports_the_oracle():
    for f in (ROOT / "hicdiff_b200").rglob("*.py"):
        src = f.read_text()
        assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), f"{f} imports the oracle"


def _header_symbols():
    hdr = (ROOT / "include" / "hicdiff_b200.h").read_text()
    return sorted(set(re.findall(r"HD_API\s+[\w\s\*]+?\b(hd_\w+)\s*\(", hdr)))


def test_c_abi_library_loads_and_exports_every_declared_symbol():
    from hicdiff_b200 import _lib

    syms = _header_symbols()
    assert len(syms) >= 24 and "hd_sample" in syms and "hd_tile_scatter" in syms
    assert sorted(_lib.SIGNATURES) == syms, "ctypes table and header disagree"
    lib = _lib.load()
    exported = subprocess.run(["nm", "-D", "--defined-only", str(_lib.lib_path())], capture_output=True, text=True).stdout
    for s in syms:
        assert getattr(lib, s) is not None
        assert re.search(rf"\bT {s}\b", exported), f"{s} not exported"
    assert lib.hd_abi_version() == _lib.HD_ABI_VERSION
    # host-only entry points behave without a GPU
    from oracle import hicdiff_oracle as O
    import numpy as np

    for n in (0, 1, 34, 64, 65, 588, 703, 802, 6400):
        for res in (40000, 10000):
            assert lib.hd_tile_count(n, 64, O.band_blocks_for(res)) == O.split_pieces(np.zeros((n, n), np.float32), 64, res).shape[0]
    assert lib.hd_tile_count(-1, 64, 4) == -1
    assert lib.hd_plan_create(None, None) != 0 and b"null" in lib.hd_last_error()
    cfg = _lib.hd_config()
    cfg.abi_version = 999
    h = ctypes.c_void_p()
    assert lib.hd_plan_create(ctypes.byref(cfg), ctypes.byref(h)) != 0 and b"ABI" in lib.hd_last_error()
    # argument validation of the optimiser / operator entry points happens before any CUDA call: exercisable without a GPU
    assert lib.hd_adam_create(None, None, -1, ctypes.byref(h)) != 0 and b"hd_adam_create" in lib.hd_last_error()
    assert lib.hd_adam_create(None, None, 0, None) != 0 and b"null out" in lib.hd_last_error()
    assert lib.hd_adam_step(None, None, 1e-3, 0.9, 0.999, 1e-8, 0.0, None) != 0 and b"null optimiser" in lib.hd_last_error()
    assert lib.hd_adam_state(None, 0, None, None, None) != 0
    assert lib.hd_adam_set_step(None, 3) != 0
    lib.hd_adam_destroy(None)                                  # a null handle is a no-op
    assert lib.hd_add_noise(None, None, 0.1, -1, None, None) != 0 and b"hd_add_noise" in lib.hd_last_error()
    assert lib.hd_tile_count(-1, 64, 4) == -1


def test_shard_range_partitions_exactly():
    from hicdiff_b200.shard import shard_range

    for n in (0, 1, 7, 221, 4096):
        for world in (1, 2, 3, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [e - s for s, e in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(5, 2, 2)


def _gloo_worker(rank, world, port, n_tiles, q):
    import os

    import torch.distributed as dist

    from hicdiff_b200.shard import gather_tiles, shard_range

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(5)
        full = torch.randn(n_tiles, 1, 64, 64, generator=g)
        s, e = shard_range(n_tiles, rank, world)
        got = gather_tiles(full[s:e].clone(), n_tiles)
        q.put((rank, bool(torch.equal(got, full))))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_tiles", [221, 3, 1])
def test_gather_tiles_world2_gloo(n_tiles):
    """N>1 path on CPU: ragged shards (221 = 111 + 110; 1 tile on 2 ranks) gathered back bit-exactly in global order."""
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29650 + n_tiles % 50
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, n_tiles, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(res) == [(0, True), (1, True)]


def test_drosophila_like_tile_budget():
    """BASELINE config 4: six chromosomes N = {588, 632, 703, 802, 34, 588} at 40 kb -> 221 tiles."""
    from hicdiff_b200 import _lib

    lib = _lib.load()
    assert sum(lib.hd_tile_count(n, 64, 4) for n in (588, 632, 703, 802, 34, 588)) == 221


def _allreduce_worker(rank, world, port, q):
    import torch.distributed as dist

    from hicdiff_b200 import train as T

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        flat = torch.arange(10, dtype=torch.float32) * (rank + 1)          # what a trainer's flat gradient buffer would hold
        T.allreduce_mean_(flat)
        q.put((rank, flat.tolist()))
    finally:
        dist.destroy_process_group()


def test_gradient_allreduce_mean_world2_gloo():
    """The data-parallel exchange of the training step (BASELINE config 5): ONE all-reduce of the flat gradient buffer, averaged."""
    import socket
    import torch.multiprocessing as mp

    from hicdiff_b200 import train as T

    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_allreduce_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = [i * 1.5 for i in range(10)]                                       # mean of 1x and 2x
    assert got[0] == want and got[1] == want
    solo = torch.ones(4)
    assert T.allreduce_mean_(solo) is solo and solo.tolist() == [1.0] * 4     # no process group: a no-op


def test_fused_adam_has_no_cpu_path():
    """hicdiff_b200.optim.Adam mirrors torch.optim.Adam's argument checks and refuses CPU parameters (no fallback)."""
    from hicdiff_b200.optim import Adam

    p = torch.nn.Parameter(torch.zeros(4))
    p.grad = torch.ones(4)
    for bad in (dict(lr=-1.0), dict(eps=-1.0), dict(betas=(1.0, 0.9)), dict(betas=(0.9, 1.0)), dict(weight_decay=-1.0)):
        with pytest.raises(ValueError):
            Adam([p], **bad)
    opt = Adam([p], lr=1e-3)
    assert opt.param_groups[0]["lr"] == 1e-3 and opt.param_groups[0]["betas"] == (0.9, 0.999)
    with pytest.raises(RuntimeError, match="no CPU path"):
        opt.step()
    assert torch.equal(p.detach(), torch.zeros(4))


def test_percentile_plan_matches_numpy():
    """the 'linear' percentile rule is restated locally (no private numpy helpers); ranks + weight + lerp
    reproduce np.percentile of the installed numpy bit for bit, dtype included."""
    import numpy as np

    from hicdiff_b200.prepare import percentile_lerp, percentile_plan

    rng = np.random.default_rng(0)
    for trial in range(1500):
        n = int(rng.integers(1, 4000))
        a = rng.standard_normal(n).astype(np.float32)
        if trial % 5 == 0:
            a = np.round(a, 1)                      # ties
        q = float(rng.choice([99.0, 99.9, 99.99, 50.0, 0.0, 100.0, rng.uniform(0, 100)]))
        ref = np.percentile(a, q)
        lo, hi, gamma = percentile_plan(n, q)
        srt = np.sort(a)
        got = percentile_lerp(srt[lo], srt[hi], gamma)
        assert ref == got and ref.dtype == got.dtype, (n, q, ref, got)


def test_add_noise_default_seed_is_fresh_per_call():
    """`seed=None` draws a new Philox key from torch's generator on every call (the reference draws a
    fresh randn_like per chromosome, PrepareData_linear.py:203-204); torch.manual_seed makes the sequence reproducible."""
    from hicdiff_b200 import prepare

    torch.manual_seed(5)
    a = [prepare._fresh_seed() for _ in range(3)]
    torch.manual_seed(5)
    b = [prepare._fresh_seed() for _ in range(3)]
    assert a == b and len(set(a)) == 3
    import inspect

    assert inspect.signature(prepare.add_noise).parameters["seed"].default is None
    assert inspect.signature(prepare.make_splits).parameters["seed"].default is None


def test_gradient_bucket_layout_partitions_the_parameters():
    """the Unet trainer's flat gradient buffer is laid out bucket by bucket in backward-completion order
    (hicdiff_b200/train.py).  Pure host logic: every parameter lands in exactly one bucket, buckets are contiguous and cover the
    buffer, everything that depends on the time embedding is in the last ('late') bucket, and the C side's boundaries
    (BUCKET_BOUNDARIES) are the first modules of the following bucket in backward order."""
    from hicdiff_b200 import hicdiff_condition, hicdiff_sr3
    from hicdiff_b200 import train as T

    for net in (hicdiff_condition.Unet(dim=64, dim_mults=(1, 2, 4, 8), self_condition=True),
                hicdiff_sr3.Unet(dim=64, dim_mults=(1, 2, 4, 8), self_condition=True, noise_level_emb=True)):
        names = [k for k, _ in net.named_parameters()]
        numels = [p.numel() for _, p in net.named_parameters()]
        order, ranges = T.grad_bucket_layout(names, numels)
        assert sorted(order) == list(range(len(names)))
        assert ranges[0][0] == 0 and ranges[-1][1] == sum(numels)
        assert all(a[1] == b[0] for a, b in zip(ranges, ranges[1:])) and all(hi > lo for lo, hi in ranges)
        assert len(ranges) == len(T.BUCKET_BOUNDARIES) + 1
        late = len(ranges) - 1
        for n in names:
            b = T.grad_bucket_of(n)
            if ".mlp." in n or "noise_func" in n or n.startswith(("time_mlp.", "init_conv.")):
                assert b == late, n
            if n.startswith(("final_res_block.block", "ups.3.0.block", "final_conv")):
                assert b == 0, n
            if n.startswith(("downs.0.", "downs.1.", "downs.2.")):
                assert b == late, n
        # buckets appear in the flat buffer in the order the backward completes them
        pos = {i: k for k, i in enumerate(order)}
        first_of = [min(pos[i] for i in range(len(names)) if T.grad_bucket_of(names[i]) == b) for b in range(len(ranges))]
        assert first_of == sorted(first_of)
