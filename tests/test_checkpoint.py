"""CPU: checkpoint loader mirror (nsr/train_util_diffusion.py:780-849): .pt and .safetensors round trips, key-AND-shape
matching with the reference's report of ignored tensors, strict load of the merged dict, rank-0 broadcast (gloo)."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _model(seed):
    from ln3diff_b200.dit.dit_models_xformers import TextCondDiTBlock
    from ln3diff_b200.dit.dit_trilatent import DiT_TriLatent
    torch.manual_seed(seed)
    return DiT_TriLatent(input_size=32, patch_size=2, in_channels=4, hidden_size=128, depth=2, num_heads=2, num_classes=0,
                         learn_sigma=False, context_dim=64, roll_out=True, vit_blk=TextCondDiTBlock)


@pytest.mark.parametrize("ext", [".pt", ".safetensors"])
def test_roundtrip_and_key_shape_matching(tmp_path, ext, capsys):
    from ln3diff_b200 import checkpoint as ck
    src, dst = _model(1), _model(2)
    sd = {k: v.clone() for k, v in src.state_dict().items()}
    sd["blocks.0.attn.qkv.weight"] = torch.zeros(5, 7)                  # same key, wrong shape -> ignored
    sd["not.in.model"] = torch.ones(3)                                   # unknown key -> ignored
    del sd["final_layer.linear.bias"]                                    # absent -> the model keeps its own
    path = ck.save_checkpoint(sd, str(tmp_path / ("w" + ext)))
    before = {k: v.clone() for k, v in dst.state_dict().items()}
    rep = ck.load_matching(dst, path)
    out = capsys.readouterr().out
    assert "!!!! ignore key:  blocks.0.attn.qkv.weight" in out and "shape in model: (384, 128)" in out
    assert "not.in.model not in model" in out
    after = dst.state_dict()
    assert torch.equal(after["blocks.1.attn.qkv.weight"], src.state_dict()["blocks.1.attn.qkv.weight"])
    assert torch.equal(after["blocks.0.attn.qkv.weight"], before["blocks.0.attn.qkv.weight"])      # mismatch kept
    assert torch.equal(after["final_layer.linear.bias"], before["final_layer.linear.bias"])        # missing kept
    assert sorted(k for k, *_ in rep.ignored) == ["blocks.0.attn.qkv.weight", "not.in.model"]
    assert "final_layer.linear.bias" in rep.kept and len(rep.loaded) == len(src.state_dict()) - 2


def test_missing_file_raises(tmp_path):
    from ln3diff_b200 import checkpoint as ck
    with pytest.raises(FileNotFoundError):
        ck.read_state_dict(str(tmp_path / "nope.pt"))


def _sync_worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from ln3diff_b200 import checkpoint as ck
    m = _model(10 + rank)                                  # different weights per rank before the sync
    ck.load_matching(m, {} if rank else {k: v + 1 for k, v in _model(10).state_dict().items()}, verbose=False)
    ret[rank] = float(sum(v.double().sum() for v in m.state_dict().values()))
    dist.barrier()
    dist.destroy_process_group()


def test_rank0_broadcast_two_ranks():
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_sync_worker, args=(2, 33500 + os.getpid() % 2000, ret), nprocs=2, join=True)
    assert ret[0] == ret[1]
