"""Route the reference's own import names to this package -- the drop-in switch.

The reference scripts (`scripts/vit_triplane_diffusion_sample*.py`) run from the repository root with
`sys.path.append('.')` and import its packages by their top-level names (`dit.dit_trilatent`,
`sgm.modules.diffusionmodules.sampling`, `nsr.volumetric_rendering.renderer`, ...).  `install()` puts an
import hook in front of the normal path finders that answers exactly the module names this package
mirrors (MIRRORED below) with the `ln3diff_b200.*` implementation and leaves every other name --
parent packages, datasets, conditioners, training utilities -- to the reference checkout on `sys.path`:

    import ln3diff_b200.overlay as overlay
    overlay.install()                      # before the reference modules are imported
    from dit.dit_trilatent import DiT_models        # -> ln3diff_b200.dit.dit_trilatent
    from nsr.train_util_diffusion import ...        # -> the untouched reference file

`uninstall()` removes the hook and the aliases it created.  Nothing is copied or patched on disk.
"""
from __future__ import annotations

import importlib
import importlib.abc
import importlib.util
import sys

# reference module name -> mirror (one line per file of the hot path)
MIRRORED = {
    "dit.dit_trilatent": "ln3diff_b200.dit.dit_trilatent",
    "dit.dit_i23d": "ln3diff_b200.dit.dit_i23d",
    "dit.dit_decoder": "ln3diff_b200.dit.dit_decoder",
    "dit.dit_models_xformers": "ln3diff_b200.dit.dit_models_xformers",
    "sgm.modules.diffusionmodules.sampling": "ln3diff_b200.sgm.modules.diffusionmodules.sampling",
    "sgm.modules.diffusionmodules.denoiser": "ln3diff_b200.sgm.modules.diffusionmodules.denoiser",
    "sgm.modules.diffusionmodules.denoiser_scaling": "ln3diff_b200.sgm.modules.diffusionmodules.denoiser_scaling",
    "sgm.modules.diffusionmodules.discretizer": "ln3diff_b200.sgm.modules.diffusionmodules.discretizer",
    "sgm.modules.diffusionmodules.guiders": "ln3diff_b200.sgm.modules.diffusionmodules.guiders",
    "guided_diffusion.gaussian_diffusion": "ln3diff_b200.guided_diffusion.gaussian_diffusion",
    "guided_diffusion.respace": "ln3diff_b200.guided_diffusion.respace",
    "transport": "ln3diff_b200.transport",
    "transport.transport": "ln3diff_b200.transport.transport",
    "nsr.triplane": "ln3diff_b200.nsr.triplane",
    "nsr.volumetric_rendering.renderer": "ln3diff_b200.nsr.volumetric_rendering.renderer",
    "nsr.volumetric_rendering.ray_sampler": "ln3diff_b200.nsr.volumetric_rendering.ray_sampler",
    "vit.vit_triplane": "ln3diff_b200.vit.vit_triplane",
    "ldm.modules.diffusionmodules.model": "ln3diff_b200.ldm.modules.diffusionmodules.model",
}


class _AliasLoader(importlib.abc.Loader):
    def __init__(self, target: str):
        self.target = target

    def create_module(self, spec):
        return importlib.import_module(self.target)      # the mirror module object itself

    def exec_module(self, module):                       # already executed under its own name
        pass


class _MirrorFinder(importlib.abc.MetaPathFinder):
    def find_spec(self, fullname, path=None, target=None):
        tgt = MIRRORED.get(fullname)
        if tgt is None:
            return None
        is_pkg = tgt == "ln3diff_b200.transport"
        spec = importlib.util.spec_from_loader(fullname, _AliasLoader(tgt), is_package=is_pkg)
        if is_pkg:
            spec.submodule_search_locations = list(importlib.import_module(tgt).__path__)
        return spec


_finder: _MirrorFinder | None = None


def install() -> None:
    """Install the import hook (idempotent).  Reference modules of MIRRORED that were imported before this
    call stay what they were: call it first."""
    global _finder
    if _finder is None:
        _finder = _MirrorFinder()
        sys.meta_path.insert(0, _finder)


def uninstall() -> None:
    global _finder
    if _finder is not None:
        sys.meta_path.remove(_finder)
        _finder = None
    for name, tgt in MIRRORED.items():
        m = sys.modules.get(name)
        if m is not None and getattr(m, "__name__", None) == tgt:
            del sys.modules[name]
