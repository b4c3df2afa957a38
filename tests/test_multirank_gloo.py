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
