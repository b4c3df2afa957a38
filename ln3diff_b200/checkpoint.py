"""Checkpoint loading / saving for the mirrored modules.

Mirror of `TrainLoop._load_and_sync_parameters` and `save` (nsr/train_util_diffusion.py:755-849):
  * `.pt` files (`torch.load`, the reference's `dist_util.load_state_dict`) and `.safetensors` files
    (`safetensors.torch.load_file`; the release weights `yslan/LN3Diff/*.safetensors` are fetched with
    `hf_hub_download` when a repo id is given and the hub is reachable);
  * a tensor is copied only when its KEY exists in the model AND its SHAPE matches (:814-817); everything else
    is reported -- "ignore key" with the checkpoint shape, and the model's shape or "not in model" (:819-830) --
    and the model keeps its own value; the merged dict is then loaded with strict=True (:832);
  * with more than one rank the parameters are broadcast from rank 0 (`dist_util.sync_params`, :838-841).
The mirrored modules keep the reference's state_dict keys and shapes, so release
checkpoints load without a key map; `load_state_dict` on a mirror invalidates its bf16 repacks and CUDA graphs.
"""
from __future__ import annotations

import os
from dataclasses import dataclass, field

import torch

HF_REPO = "yslan/LN3Diff"      # nsr/train_util_diffusion.py:808


@dataclass
class LoadReport:
    loaded: list = field(default_factory=list)            # keys copied from the checkpoint
    ignored: list = field(default_factory=list)           # (key, checkpoint shape, model shape or None)
    kept: list = field(default_factory=list)              # model keys the checkpoint did not provide

    def __str__(self):
        lines = [f"loaded {len(self.loaded)} tensors, ignored {len(self.ignored)}, kept {len(self.kept)} model tensors"]
        for k, cs, ms in self.ignored:
            lines.append(f"!!!! ignore key:  {k} :  {tuple(cs)}" + (f"  shape in model: {tuple(ms)}" if ms is not None
                                                                     else f"  {k} not in model"))
        return "\n".join(lines)


def read_state_dict(path: str, map_location="cpu", hf_repo: str | None = None) -> dict:
    """`.safetensors` -> safetensors.torch.load_file; anything else -> torch.load (weights only).  A path that does
    not exist locally is looked up in `hf_repo` (default: the release repo for *.safetensors names, as the reference
    does) through huggingface_hub -- which needs network access and raises otherwise."""
    if not os.path.exists(path):
        repo = hf_repo or (HF_REPO if path.endswith(".safetensors") else None)
        if repo is None:
            raise FileNotFoundError(path)
        try:
            from huggingface_hub import hf_hub_download
            path = hf_hub_download(repo_id=repo, filename=path)
        except Exception as e:
            raise FileNotFoundError(f"{path}: not a local file and could not be fetched from {repo}: {e}") from e
    if path.endswith(".safetensors"):
        from safetensors.torch import load_file
        dev = map_location if isinstance(map_location, str) else str(map_location)
        return load_file(path, device=dev)
    sd = torch.load(path, map_location=map_location, weights_only=True)
    if isinstance(sd, dict) and "state_dict" in sd and all(isinstance(v, torch.Tensor) for v in sd["state_dict"].values()):
        sd = sd["state_dict"]
    return sd


@torch.no_grad()
def load_matching(model: torch.nn.Module, checkpoint, *, map_location="cpu", verbose: bool = True,
                  sync: bool = True, group=None) -> LoadReport:
    """Copy every tensor of `checkpoint` (a path or a state_dict) whose key and shape match into `model`
    (reference :811-832), report the rest, then broadcast from rank 0 when torch.distributed is initialised."""
    resume = read_state_dict(checkpoint, map_location) if isinstance(checkpoint, (str, os.PathLike)) else checkpoint
    target = model.state_dict()
    rep = LoadReport()
    for k, v in resume.items():
        if k in target and tuple(v.shape) == tuple(target[k].shape):
            target[k] = v
            rep.loaded.append(k)
        else:
            rep.ignored.append((k, tuple(v.shape), tuple(target[k].shape) if k in target else None))
    rep.kept = [k for k in target if k not in set(rep.loaded)]
    model.load_state_dict(target, strict=True)
    if verbose and rep.ignored:
        print(rep)
    if sync:
        sync_params(model, group=group)
    return rep


@torch.no_grad()
def sync_params(model: torch.nn.Module, src: int = 0, group=None) -> None:
    """dist_util.sync_params: broadcast every parameter and buffer from `src` (no-op on one rank)."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return
    for t in list(model.parameters()) + list(model.buffers()):
        dist.broadcast(t.data, src=src, group=group)
    if hasattr(model, "_invalidate"):
        model._invalidate()          # derived bf16 repacks / graphs must be rebuilt from the synced weights


def save_checkpoint(model_or_state: torch.nn.Module | dict, path: str) -> str:
    """reference :755-776 (`th.save(state_dict, f)`), plus `.safetensors` output for the release format."""
    sd = model_or_state.state_dict() if isinstance(model_or_state, torch.nn.Module) else model_or_state
    sd = {k: v.detach().cpu().contiguous() for k, v in sd.items()}
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    if path.endswith(".safetensors"):
        from safetensors.torch import save_file
        save_file(sd, path)
    else:
        torch.save(sd, path)
    return path
