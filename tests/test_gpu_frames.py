"""GPU (-m gpu): the frame sink (`ln3_pack_frames`, byte-exact against the numpy restatement of the
reference's per-view host loop, nsr/train_util_diffusion.py:292-376) and the prompt-sharded end-to-end entry
point `pipeline.generate_sharded` on one rank."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available(), "-m gpu tests need a GPU"
    from ln3diff_b200 import _lib
    _lib.lib()
    return torch.device("cuda", 0)


@pytest.mark.parametrize("N,H,W", [(1, 8, 8), (3, 32, 48), (5, 128, 128), (2, 256, 256)])
def test_pack_frames_bit_exact(dev, N, H, W):
    from ln3diff_b200.frames import FrameSink, viridis_table
    from oracle.frames import video_frames
    g = torch.Generator().manual_seed(N * H + W)
    img = torch.rand(N, 3, H, W, generator=g) * 2.4 - 1.2              # beyond [-1, 1]: both clips are exercised
    img[0, :, 0, :4] = torch.tensor([-1.0, 1.0, 0.0, 2.0 / 255 - 1])    # exact byte boundaries
    dep = torch.rand(N, 1, H, W, generator=g) * 1.3 + 1.1
    if N > 1:
        dep[1] = 1.7                                                    # constant view: max == min -> "bad" colour
    sink = FrameSink(dev)
    table = viridis_table()
    both = sink.pack(img.to(dev), dep.to(dev))
    assert both.dtype == torch.uint8 and both.shape == (N, H, 2 * W, 3)
    ref = video_frames(img.numpy(), dep.numpy(), table)
    assert np.array_equal(both.cpu().numpy(), ref)
    rgb = sink.pack(img.to(dev))
    assert rgb.shape == (N, H, W, 3)
    assert np.array_equal(rgb.cpu().numpy(), video_frames(img.numpy(), None, table))
    host = sink.to_host(both)
    torch.cuda.synchronize()
    assert host.is_pinned() and np.array_equal(host.numpy(), ref)


def test_pack_frames_leading_dims_and_custom_colormap(dev):
    from ln3diff_b200.frames import FrameSink
    from oracle.frames import video_frames
    rng = np.random.default_rng(0)
    table = rng.uniform(0, 1, (256, 3))
    g = torch.Generator().manual_seed(9)
    img = torch.rand(2, 3, 3, 16, 16, generator=g) * 2 - 1
    dep = torch.rand(2, 3, 1, 16, 16, generator=g) + 1
    out = FrameSink(dev, colormap=table).pack(img.to(dev), dep.to(dev))
    assert out.shape == (2, 3, 16, 32, 3)
    ref = video_frames(img.reshape(6, 3, 16, 16).numpy(), dep.reshape(6, 1, 16, 16).numpy(), table)
    assert np.array_equal(out.cpu().numpy().reshape(6, 16, 32, 3), ref)


def test_pack_frames_rejects_bad_input(dev):
    from ln3diff_b200 import ops
    with pytest.raises(ValueError, match="multiple of 4"):
        ops.pack_frames(torch.zeros(1, 3, 8, 6, device=dev))
    with pytest.raises(ValueError, match="CUDA"):
        ops.pack_frames(torch.zeros(1, 3, 8, 8))


def test_generate_sharded_single_rank_matches_stagewise(dev):
    """generate_sharded on one rank == its three stages called by hand with the same batch structure and RNG
    state: global CPU noise draw, sample_t23d, decode_and_render, frame sink (3 prompts in batches of 2 + 1)."""
    from ln3diff_b200 import pipeline
    from ln3diff_b200.frames import FrameSink
    from ln3diff_b200.utils import build_ae_decoder, build_t23d, orbit_cameras
    m = build_t23d("DiT-B/2", device=dev)
    dec = build_ae_decoder("DiT2-S/2", device=dev)
    cams = orbit_cameras(2)
    g = torch.Generator().manual_seed(8)
    P = 3
    c_all = {"crossattn": torch.randn(P, 77, 768, generator=g)}
    uc_all = {"crossattn": torch.zeros(P, 77, 768)}
    torch.manual_seed(0)
    out = pipeline.generate_sharded(m, dec, c_all, uc_all, cams, seed=41, num_steps=4, scale=6.5, resolution=32, batch=2,
                                    with_depth=True)
    assert out["frames_all"].shape == (P, 2, 32, 64, 3) and out["shard"] == (0, P)
    randn_all = torch.randn(P, 12, 32, 32, generator=torch.Generator().manual_seed(41))
    sink = FrameSink(dev)
    torch.manual_seed(0)
    frames, lats = [], []
    for sl in (slice(0, 2), slice(2, 3)):
        lat = pipeline.sample_t23d(m, randn_all[sl].to(dev), {"crossattn": c_all["crossattn"][sl].to(dev)},
                                   {"crossattn": uc_all["crossattn"][sl].to(dev)}, 4, 6.5)
        r = pipeline.decode_and_render(dec, lat, cams, 32)
        frames.append(sink.pack(r["image_raw"], r["image_depth"]))
        lats.append(lat)
    assert torch.equal(out["latents"], torch.cat(lats))
    assert torch.equal(out["frames_all"], torch.cat(frames))
    assert out["frames_all"].float().std() > 1.0                      # not a blank video
