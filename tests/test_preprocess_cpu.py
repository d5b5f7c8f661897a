"""CPU tests of the preprocessing row (no GPU): the oracle's restatement of Pillow's uint8 bicubic resample is bit-exact
against Pillow itself, and the product's vectorised coefficient tables (actionmesh_b200/preprocess.py, what the CUDA
kernels consume) are identical to the oracle's scalar restatement and reproduce Pillow when driven through an integer
convolution."""
import numpy as np
import pytest
from PIL import Image

from actionmesh_b200 import preprocess as pp
from oracle import preprocess_oracle as po

SIZES = [(512, 512, 256), (300, 400, 256), (224, 224, 256), (100, 130, 256), (720, 1280, 256), (256, 256, 256), (257, 300, 256),
         (1080, 607, 256)]


@pytest.mark.parametrize("h,w,s", SIZES)
def test_oracle_resize_is_bit_exact_vs_pillow(h, w, s):
    rng = np.random.default_rng(h * 1000 + w)
    a = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    oh, ow = po.resize_output_size(h, w, s)
    ref = np.asarray(Image.fromarray(a, "RGB").resize((ow, oh), resample=Image.BICUBIC))
    assert np.array_equal(po.pil_bicubic_resize_u8(a, oh, ow), ref)


@pytest.mark.parametrize("n_in,n_out", [(512, 256), (400, 341), (224, 256), (130, 332), (1280, 455), (300, 298), (607, 256),
                                        (1080, 455), (37, 256), (4000, 256)])
def test_product_tables_equal_oracle_tables(n_in, n_out):
    b, k = pp.resample_table(n_in, n_out)
    ob, ok = po._coeffs(n_in, n_out)
    assert np.array_equal(b, np.array(ob, dtype=np.int32))
    assert np.array_equal(k, np.array(ok, dtype=np.int32))


def test_product_plan_reproduces_the_reference_pipeline():
    """Drive the product's cropped tables through a numpy integer convolution (what the two kernels do) and compare with
    the reference pipeline restated on Pillow: uint8 crop bit-identical, float output identical."""
    rng = np.random.default_rng(3)
    for (h, w) in [(512, 512), (300, 400), (1080, 607), (256, 300)]:
        a = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        ref_pv, ref_u8 = po.bit_preprocess_pil([Image.fromarray(a, "RGB")], return_u8=True)
        oh, ow = pp.resize_output_size(h, w, 256)
        assert (oh, ow) == po.resize_output_size(h, w, 256)
        top, left = (oh - 224) // 2, (ow - 224) // 2
        bh, kh = (pp.resample_table(w, ow) if w != ow else (None, None))
        bv, kv = (pp.resample_table(h, oh) if h != oh else (None, None))
        x = a.astype(np.int64)
        if bh is not None:
            mid = np.empty((h, 224, 3), dtype=np.int64)
            for j in range(224):
                xmin, n = bh[left + j]
                acc = (1 << 21) + (x[:, xmin:xmin + n] * kh[left + j, :n].astype(np.int64)[None, :, None]).sum(1)
                mid[:, j] = np.clip(acc >> 22, 0, 255)
        else:
            mid = x[:, left:left + 224]
        if bv is not None:
            out = np.empty((224, 224, 3), dtype=np.int64)
            for i in range(224):
                ymin, n = bv[top + i]
                acc = (1 << 21) + (mid[ymin:ymin + n] * kv[top + i, :n].astype(np.int64)[:, None, None]).sum(0)
                out[i] = np.clip(acc >> 22, 0, 255)
        else:
            out = mid[top:top + 224]
        assert np.array_equal(out.astype(np.uint8), ref_u8[0])
        p = pp.B200ImagePreprocessor()
        pv = (p._lut_host[out] - np.array(p.image_mean, dtype=np.float32)) / np.array(p.image_std, dtype=np.float32)
        assert np.array_equal(np.transpose(pv.astype(np.float32), (2, 0, 1)), ref_pv[0])


def test_preprocessor_refuses_cpu():
    import torch

    from actionmesh_b200._lib import AmbError

    with pytest.raises(AmbError):
        pp.B200ImagePreprocessor().preprocess_u8(torch.zeros(1, 32, 32, 3, dtype=torch.uint8), "cpu")


def test_random_sizes_against_pillow_property():
    """Property test over random (in, out) sizes incl. up-scaling, extreme aspect ratios and tiny images: the product table
    driven through the integer convolution equals Pillow on one axis (a 1-row / 1-column strip keeps it fast)."""
    from hypothesis import given, settings
    from hypothesis import strategies as st

    @settings(max_examples=60, deadline=None)
    @given(n_in=st.integers(2, 1500), n_out=st.integers(2, 700), seed=st.integers(0, 2 ** 16))
    def check(n_in, n_out, seed):
        rng = np.random.default_rng(seed)
        strip = rng.integers(0, 256, (3, n_in, 3), dtype=np.uint8)            # 3 rows, n_in columns
        ref = np.asarray(Image.fromarray(strip, "RGB").resize((n_out, 3), resample=Image.BICUBIC))
        if n_in == n_out:
            assert np.array_equal(ref, strip)
            return
        b, k = pp.resample_table(n_in, n_out)
        x = strip.astype(np.int64)
        out = np.empty((3, n_out, 3), dtype=np.int64)
        for j in range(n_out):
            xmin, n = b[j]
            acc = (1 << 21) + (x[:, xmin:xmin + n] * k[j, :n].astype(np.int64)[None, :, None]).sum(1)
            out[:, j] = np.clip(acc >> 22, 0, 255)
        assert np.array_equal(out.astype(np.uint8), ref), (n_in, n_out)

    check()


def test_c_oracle_is_bit_exact_vs_pillow():
    """oracle/pil_resample.c (plain-C restatement of Pillow's Resample.c, built by __graft_entry__.build()) against Pillow."""
    import ctypes
    import os
    import subprocess

    from conftest import ROOT

    so = os.path.join(ROOT, "oracle", "_build", "libpil_resample.so")
    if not os.path.exists(so):
        subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle")], check=True)
    lib = ctypes.CDLL(so)
    lib.amb_oracle_resize_u8.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                         ctypes.c_void_p]
    rng = np.random.default_rng(11)
    for (h, w, oh, ow, c) in [(512, 512, 256, 256, 3), (300, 400, 256, 341, 3), (96, 130, 256, 346, 3), (256, 300, 256, 300, 3),
                              (257, 256, 257, 256, 3), (720, 405, 455, 256, 3)]:   # RGB only: PIL premultiplies RGBA; the
                                                                                  # pipeline converts to RGB first
        a = np.ascontiguousarray(rng.integers(0, 256, (h, w, c), dtype=np.uint8))
        out = np.empty((oh, ow, c), dtype=np.uint8)
        assert lib.amb_oracle_resize_u8(a.ctypes.data, h, w, c, oh, ow, out.ctypes.data) == 0
        ref = np.asarray(Image.fromarray(a, "RGB").resize((ow, oh), resample=Image.BICUBIC))
        assert np.array_equal(out, ref), (h, w, oh, ow, c)
