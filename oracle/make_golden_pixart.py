"""TEST INFRASTRUCTURE.  Golden vector for the PixArt-style T23D denoiser, produced by the reference's
own `DiT_TriLatent_PixelArt` (dit/dit_trilatent.py:146-246, registry key 'DiT-PixelArt-B/2') run in the
build container on the seeded inputs / weights of oracle.fixtures.

Run:  python oracle/make_golden_pixart.py     (needs /root/reference; writes tests/golden/dit_t23d_pixart.npz)
"""
import contextlib
import io
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import _stubs  # noqa: E402

_stubs.install()
sys.path.insert(0, "/root/reference")
_stubs.patch_dit_namespace()
from oracle import fixtures as fx  # noqa: E402


def main():
    import dit.dit_trilatent as dt
    with contextlib.redirect_stdout(io.StringIO()):
        ref = dt.DiT_models[fx.T23D_PIXART_ARCH](input_size=32, num_classes=0, learn_sigma=False, in_channels=4,
                                                 context_dim=768, roll_out=True)
    ref.eval()
    assert type(ref).__name__ == "DiT_TriLatent_PixelArt" and type(ref.blocks[0]).__name__ == "PixelArtTextCondDiTBlock"
    shapes = {k: tuple(v.shape) for k, v in ref.state_dict().items()}
    ref.load_state_dict(fx.i23d_state_dict(shapes, ref.state_dict()["pos_embed"]))
    x, t, ctx = fx.t23d_pixart_inputs()
    with torch.no_grad():
        y = ref(x, t, ctx)
        yc = ref.forward_with_cfg(x, t, ctx, 6.5)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "dit_t23d_pixart.npz"), out=y.numpy(), out_cfg=yc.numpy())
    print("dit_t23d_pixart", y.shape, float(y.abs().max()))


if __name__ == "__main__":
    main()
