"""CPU, world_size 2 (gloo): the N>1 host logic of the prompt-sharded pipeline -- every rank draws
the same global noise (one manual_seed + one randn, as the reference engine does,
nsr/lsgm/sgm_DiffusionEngine.py:457,395), takes its contiguous slice, and one all_gather reassembles
results in prompt order.  The 'model' is the closed-form toy network so the test runs without a GPU."""
import os

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import fixtures as fx
    from oracle import samplers as osmp
    toy = fx.toy_network()
    per = 2
    g = torch.Generator().manual_seed(41)
    randn_all = torch.randn(per * world, 12, 32, 32, generator=g)
    ctx_all = torch.randn(per * world, 77, 16, generator=g)
    sl = slice(rank * per, (rank + 1) * per)
    c = {"crossattn": ctx_all[sl]}
    uc = {"crossattn": torch.zeros_like(ctx_all[sl])}
    local = osmp.euler_edm_cfg_sample(toy, randn_all[sl].clone(), c, uc, 5, 6.5)
    gathered = torch.empty(per * world, 12, 32, 32)
    dist.all_gather_into_tensor(gathered, local.contiguous())
    # timing protocol: max over ranks
    t = torch.tensor([float(rank + 1)])
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        ret["gathered"] = gathered
        ret["tmax"] = t.item()
    dist.barrier()
    dist.destroy_process_group()


def test_prompt_sharding_two_ranks_matches_single_process():
    from oracle import fixtures as fx
    from oracle import samplers as osmp
    world, per = 2, 2
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    toy = fx.toy_network()
    g = torch.Generator().manual_seed(41)
    randn_all = torch.randn(per * world, 12, 32, 32, generator=g)
    ctx_all = torch.randn(per * world, 77, 16, generator=g)
    ref = osmp.euler_edm_cfg_sample(toy, randn_all.clone(), {"crossattn": ctx_all},
                                    {"crossattn": torch.zeros_like(ctx_all)}, 5, 6.5)
    assert torch.equal(ret["gathered"], ref)      # sharding is exact: samples are independent
    assert ret["tmax"] == 2.0


# ---------------------------------------------------------------- the product's sharded entry point
def _cpu_stages():
    """CPU stand-ins for the three CUDA stages of pipeline.generate_sharded (sample / decode+render / pack):
    cheap, deterministic, and per-sample independent like the real stages."""
    def sample_fn(x, c, uc):
        return x * 0.5 + c["crossattn"].mean(dim=(1, 2))[:, None, None, None] - uc["crossattn"].sum(dim=(1, 2))[:, None, None, None]

    def render_fn(lat):                       # (n,12,32,32) -> image_raw (n, V=3, 3, 8, 8), depth (n, 3, 1, 8, 8)
        n = lat.shape[0]
        base = lat[:, :3, :8, :8]
        img = torch.stack([torch.tanh(base * (v + 1)) for v in range(3)], 1)
        return dict(image_raw=img, image_depth=img[:, :, :1] + 2.0)

    def pack_fn(r):
        return ((r["image_raw"].permute(0, 1, 3, 4, 2).double() * 127.5 + 127.5).clamp(0, 255)).to(torch.uint8)
    return sample_fn, render_fn, pack_fn


def _sharded_worker(rank, world, port, P, batch, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from ln3diff_b200 import pipeline
    g = torch.Generator().manual_seed(3)
    c_all = {"crossattn": torch.randn(P, 77, 16, generator=g)}
    uc_all = {"crossattn": torch.zeros(P, 77, 16)}
    s, r, p = _cpu_stages()
    out = pipeline.generate_sharded(None, None, c_all, uc_all, torch.zeros(3, 25), seed=41, resolution=8, batch=batch,
                                    device="cpu", sample_fn=s, render_fn=r, pack_fn=p)
    ret[rank] = dict(frames_all=out["frames_all"].clone(), shard=out["shard"], latents=out["latents"].clone(),
                     bytes=out["gather_bytes_per_rank"])
    dist.barrier()
    dist.destroy_process_group()


def _run_sharded(P, batch, world=2):
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 31500 + (os.getpid() % 2000) + P
    mp.spawn(_sharded_worker, args=(world, port, P, batch, ret), nprocs=world, join=True)
    return ret


def test_generate_sharded_two_ranks_equals_one_process():
    """pipeline.generate_sharded (the product entry point, CPU stand-in stages) on 2 gloo ranks reproduces the
    single-process result: every rank slices the same global noise draw, frames come back in prompt order on
    every rank, for an even split, a ragged split (P = 5: ranks hold 3 + 2) and multi-batch shards."""
    from ln3diff_b200 import pipeline
    s, r, p = _cpu_stages()
    for P, batch in ((4, 32), (5, 32), (5, 2), (1, 32)):
        g = torch.Generator().manual_seed(3)
        c_all = {"crossattn": torch.randn(P, 77, 16, generator=g)}
        uc_all = {"crossattn": torch.zeros(P, 77, 16)}
        one = pipeline.generate_sharded(None, None, c_all, uc_all, torch.zeros(3, 25), seed=41, resolution=8,
                                        batch=batch, device="cpu", sample_fn=s, render_fn=r, pack_fn=p)
        assert one["frames_all"].shape == (P, 3, 8, 8, 3) and one["shard"] == (0, P)
        ret = _run_sharded(P, batch)
        per = -(-P // 2)
        assert ret[0]["shard"] == (0, min(per, P)) and ret[1]["shard"] == (min(per, P), P)
        for rank in (0, 1):
            assert torch.equal(ret[rank]["frames_all"], one["frames_all"]), (P, batch, rank)
            lo, hi = ret[rank]["shard"]
            assert torch.equal(ret[rank]["latents"], one["latents"][lo:hi])
            assert ret[rank]["bytes"] == per * 3 * 8 * 8 * 3


def test_shard_range_covers_every_prompt_once():
    from ln3diff_b200.pipeline import shard_range
    for P in (0, 1, 7, 8, 9, 256):
        for G in (1, 2, 4, 8):
            seen = []
            for r in range(G):
                lo, hi, per = shard_range(P, G, r)
                assert 0 <= hi - lo <= per
                seen += list(range(lo, hi))
            assert seen == list(range(P))
