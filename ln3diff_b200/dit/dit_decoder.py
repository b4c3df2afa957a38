"""Mirror of reference dit/dit_decoder.py: DiT2 (the VAE decoder backbone) + DiT2_models.

DiT2 has no x/t embedders and no final layer (dit_decoder.py:91-93); its input `c` (B, 768, D) is the
per-token adaLN condition and the stream starts from the learned pos_embed.  Parameters only; the
forward runs inside ln3diff_b200.vit.vit_triplane."""
import torch
import torch.nn as nn

from .dit_models_xformers import DiTBlock


def modulate2(x, shift, scale):
    return x * (1 + scale) + shift


class DiTBlock2(DiTBlock):
    pass


class DiT2(nn.Module):
    def __init__(self, input_size=32, patch_size=2, in_channels=4, hidden_size=1152, depth=28, num_heads=16,
                 mlp_ratio=4, class_dropout_prob=0.1, num_classes=1000, learn_sigma=True, mixing_logit_init=-3,
                 mixed_prediction=True, context_dim=False, roll_out=False, plane_n=3, return_all_layers=False,
                 vit_blk=...):
        super().__init__()
        if hidden_size // num_heads != 64:
            raise NotImplementedError("libln3b200 attention implements head_dim=64")
        if return_all_layers or not roll_out:
            raise NotImplementedError("release decoder: roll_out=True, return_all_layers=False")
        self.depth, self.embed_dim, self.num_heads, self.mlp_ratio = depth, hidden_size, num_heads, mlp_ratio
        self.plane_n, self.roll_out, self.return_all_layers = plane_n, roll_out, return_all_layers
        num_patches = (input_size // patch_size) ** 2
        self.pos_embed = nn.Parameter(torch.zeros(1, num_patches, hidden_size), requires_grad=False)
        self.blocks = nn.ModuleList([DiTBlock2(hidden_size, num_heads, mlp_ratio=mlp_ratio) for _ in range(depth)])
        self.clip_text_proj = None
        for m in self.modules():
            if isinstance(m, nn.Linear):
                nn.init.xavier_uniform_(m.weight)
                if m.bias is not None:
                    nn.init.constant_(m.bias, 0)
        for b in self.blocks:
            nn.init.constant_(b.adaLN_modulation[-1].weight, 0)
            nn.init.constant_(b.adaLN_modulation[-1].bias, 0)


def _mk(depth, hidden, heads):
    def f(**kwargs):
        return DiT2(depth=depth, hidden_size=hidden, patch_size=2, num_heads=heads, **kwargs)
    return f


# reference dit/dit_decoder.py:272-288 (patch-2 entries with head_dim 64)
DiT2_models = {"DiT2-L/2": _mk(24, 1024, 16), "DiT2-L/2-half": _mk(12, 1024, 16), "DiT2-B/2": _mk(12, 768, 12),
               "DiT2-S/2": _mk(12, 384, 6)}
