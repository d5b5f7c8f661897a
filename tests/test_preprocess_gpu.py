"""GPU parity of the image-preprocessing kernels (-m gpu): bit-exact on the resized + cropped uint8 image against Pillow
(the reference's transformers<5 BitImageProcessor path restated in oracle/preprocess_oracle.py), and identical fp32
pixel_values (same fp32 subtract/divide, IEEE round-to-nearest)."""
import numpy as np
import pytest
import torch
from PIL import Image

from oracle import preprocess_oracle as po

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("h,w,mode", [(512, 512, "RGB"), (512, 512, "RGBA"), (300, 400, "RGB"), (1080, 607, "RGB"),
                                      (224, 224, "RGB"), (256, 300, "RGB"), (257, 256, "RGBA"), (96, 130, "RGB")])
def test_preprocess_bit_exact_vs_pillow(amb_lib, h, w, mode):
    from actionmesh_b200.preprocess import B200ImagePreprocessor

    rng = np.random.default_rng(h * 7 + w)
    imgs = [Image.fromarray(rng.integers(0, 256, (h, w, len(mode)), dtype=np.uint8), mode) for _ in range(3)]
    ref_pv, ref_u8 = po.bit_preprocess_pil(imgs, return_u8=True)
    p = B200ImagePreprocessor()
    frames = torch.from_numpy(np.stack([np.asarray(im) for im in imgs]))          # RGBA frames keep their 4th byte: skipped by the kernel
    pv, u8 = p.preprocess_u8(frames, "cuda", return_u8=True)
    assert np.array_equal(u8.cpu().numpy(), ref_u8)
    assert np.array_equal(pv.cpu().numpy(), ref_pv)
    pv2 = p.preprocess(imgs, "cuda")                                              # the PIL-list entry point
    assert torch.equal(pv2, pv)


def test_mixed_sizes_and_empty_rows_edge(amb_lib):
    from actionmesh_b200.preprocess import B200ImagePreprocessor

    rng = np.random.default_rng(1)
    imgs = [Image.fromarray(rng.integers(0, 256, s + (3,), dtype=np.uint8), "RGB") for s in ((300, 400), (512, 512), (300, 400))]
    pv = B200ImagePreprocessor().preprocess(imgs, "cuda").cpu().numpy()
    assert np.array_equal(pv, po.bit_preprocess_pil(imgs))


def test_encoder_uses_gpu_preprocessing(amb_lib):
    """encode_images(list of PIL) == encode_pixel_values(reference pixel_values): the preprocessing seam is exact."""
    from actionmesh_b200.image_encoder import B200ImageEncoder

    enc = B200ImageEncoder(num_layers=2).to("cuda")
    enc.init_random_(seed=3)
    rng = np.random.default_rng(2)
    imgs = [Image.fromarray(rng.integers(0, 256, (512, 512, 4), dtype=np.uint8), "RGBA") for _ in range(2)]
    a = enc.encode_images(imgs)
    b = enc.encode_pixel_values(torch.from_numpy(po.bit_preprocess_pil(imgs)))
    assert torch.equal(a, b)
