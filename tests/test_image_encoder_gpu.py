"""DinoV2 frame encoder parity (-m gpu): B200ImageEncoder vs the reference's arithmetic (HF transformers Dinov2Model,
fp32 on the host cores).  The reference runs the encoder in fp32 (pipeline.py:664-667), and so does the default path here
(split-bf16 tensor-core GEMMs + fp32 attention): relative Frobenius error <= 1e-4 on last_hidden_state with UNROUNDED fp32
weights.  The optional bf16-operand path (precision="bf16") is held to 1e-2 with bf16-representable weights."""
import pytest
import torch

from oracle import dinov2_oracle

pytestmark = pytest.mark.gpu


def rel(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return float((a - b).norm() / b.norm())


def _pair(hidden, layers, heads, seed, precision="fp32"):
    from actionmesh_b200.image_encoder import B200ImageEncoder

    ref = dinov2_oracle.make_model(hidden, layers, heads, seed, bf16_weights=(precision == "bf16"))
    enc = B200ImageEncoder(hidden_size=hidden, num_layers=layers, num_heads=heads, precision=precision).to("cuda")
    enc.load_state_dict(ref.state_dict())
    return ref, enc


def test_small_dino_matches_hf(amb_lib):
    ref, enc = _pair(256, 2, 4, 11)
    px = torch.randn(3, 3, 224, 224, generator=torch.Generator().manual_seed(1))
    out = enc.encode_pixel_values(px)
    assert out.shape == (3, 257, 256) and out.dtype == torch.float32
    assert rel(out, dinov2_oracle.encode(ref, px)) < 1e-4


def test_dinov2_large_matches_hf(amb_lib):
    """Full DinoV2-L/14 (24 layers, width 1024), fp32-grade default path vs HF fp32: <= 1e-4."""
    ref, enc = _pair(1024, 24, 16, 1235)
    px = torch.randn(2, 3, 224, 224, generator=torch.Generator().manual_seed(2))
    out = enc.encode_pixel_values(px)
    assert out.shape == (2, 257, 1024)
    e = rel(out, dinov2_oracle.encode(ref, px))
    print("DINO_FP32_PARITY", e)
    assert e < 1e-4


def test_dinov2_large_bf16_operand_path(amb_lib):
    ref, enc = _pair(1024, 24, 16, 1235, precision="bf16")
    px = torch.randn(2, 3, 224, 224, generator=torch.Generator().manual_seed(2))
    e = rel(enc.encode_pixel_values(px), dinov2_oracle.encode(ref, px))
    print("DINO_BF16_PARITY", e)
    assert e < 1e-2


def test_attn_small_f32_matches_torch(amb_lib):
    """The fp32 short-sequence attention kernel alone vs torch fp32 SDPA math (seq 257 and a ragged 100, 3 frames, 4 heads)."""
    from actionmesh_b200 import ops

    for seq in (257, 100, 320, 1):
        T, H = 3, 4
        qkv = torch.randn(T * seq, 3 * H * 64, generator=torch.Generator().manual_seed(seq)).cuda()
        out = torch.empty(T * seq, H * 64, device="cuda")
        ops.attn_small_f32(qkv, T, seq, H, 0.125, out)
        q, k, v = (qkv[:, i * H * 64:(i + 1) * H * 64].view(T, seq, H, 64).permute(0, 2, 1, 3).double() for i in range(3))
        ref = (torch.softmax(q @ k.transpose(-1, -2) * 0.125, -1) @ v).permute(0, 2, 1, 3).reshape(T * seq, H * 64)
        assert rel(out, ref) < 2e-6, (seq, rel(out, ref))


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
    assert rel(out, dinov2_oracle.encode(ref, px)) < 1e-4
