"""DinoV2 frame encoder parity (-m gpu): B200ImageEncoder vs the reference's arithmetic (HF transformers Dinov2Model,
fp32 on the host cores) with identical bf16-representable weights.  Tolerance: fp32 residual stream, bf16 GEMM operands
and attention probabilities => relative Frobenius error <= 1e-2 on last_hidden_state."""
import pytest
import torch

from oracle import dinov2_oracle

pytestmark = pytest.mark.gpu


def rel(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return float((a - b).norm() / b.norm())


def _pair(hidden, layers, heads, seed):
    from actionmesh_b200.image_encoder import B200ImageEncoder

    ref = dinov2_oracle.make_model(hidden, layers, heads, seed)
    enc = B200ImageEncoder(hidden_size=hidden, num_layers=layers, num_heads=heads).to("cuda")
    enc.load_state_dict(ref.state_dict())
    return ref, enc


def test_small_dino_matches_hf(amb_lib):
    ref, enc = _pair(256, 2, 4, 11)
    px = torch.randn(3, 3, 224, 224, generator=torch.Generator().manual_seed(1))
    out = enc.encode_pixel_values(px)
    assert out.shape == (3, 257, 256) and out.dtype == torch.float32
    assert rel(out, dinov2_oracle.encode(ref, px)) < 1e-2


def test_dinov2_large_matches_hf(amb_lib):
    ref, enc = _pair(1024, 24, 16, 1235)
    px = torch.randn(2, 3, 224, 224, generator=torch.Generator().manual_seed(2))
    out = enc.encode_pixel_values(px)
    assert out.shape == (2, 257, 1024)
    assert rel(out, dinov2_oracle.encode(ref, px)) < 1e-2


def test_encode_images_pil_surface(amb_lib):
    """encode_images(list[PIL]) == BitImageProcessor (transformers<5 / Pillow semantics, oracle/preprocess_oracle.py) followed
    by DinoV2, like the reference (image_encoder.py:48-55); the preprocessing itself runs in CUDA (tests/test_preprocess_gpu.py)."""
    from PIL import Image
    import numpy as np

    ref, enc = _pair(256, 2, 4, 12)
    rng = np.random.default_rng(7)
    imgs = [Image.fromarray(rng.integers(0, 255, (512, 512, 3), dtype=np.uint8), "RGB") for _ in range(2)]
    out = enc.encode_images(imgs)
    from oracle import preprocess_oracle

    px = torch.from_numpy(preprocess_oracle.bit_preprocess_pil(imgs))
    assert px.shape == (2, 3, 224, 224)
    assert rel(out, dinov2_oracle.encode(ref, px)) < 1e-2
