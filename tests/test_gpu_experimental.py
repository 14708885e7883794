"""GPU parity tests of experimental::normalize_blocks: the CUDA kernels behind dltcuda_bc1_* must
reproduce the oracle's restatement of core/dxt-lossless-transform-bc1/src/experimental/normalize_blocks bit for bit."""
import numpy as np
import pytest

import dxt_lossless_transform_b200 as dlt
import oracle
from dxt_lossless_transform_b200 import experimental as ex
from dxt_lossless_transform_b200 import synth
from norm_cases import REFERENCE_VECTORS, crafted_blocks

pytestmark = pytest.mark.gpu
MODES = list(ex.ColorNormalizationMode)


def test_reference_vectors_on_the_gpu():
    for name, block, expected in REFERENCE_VECTORS:
        src = np.frombuffer(block, np.uint8).copy()
        for mode in MODES:
            out = np.zeros_like(src)
            ex.normalize_blocks(src, out, mode)
            want = block if mode == 0 or expected[int(mode)] is None else expected[int(mode)]
            assert out.tobytes() == want, (name, mode)


@pytest.mark.parametrize("n", [1, 2, 3, 255, 256, 257, 100_003])
def test_normalize_blocks_equals_oracle(n):
    for data in (crafted_blocks(n, seed=n), synth.random_blocks(1, n, seed=n), synth.texture_blocks(1, n, seed=n)):
        outs = [np.zeros_like(data) for _ in range(3)]
        any_ = ex.normalize_blocks_all_modes(data, outs)
        want, want_any = oracle.normalize_blocks_all_modes(data)
        assert any_ == want_any
        for m in range(3):
            assert np.array_equal(outs[m], want[m]), (n, m)
        for mode in MODES:
            out = np.zeros_like(data)
            ex.normalize_blocks(data, out, mode)
            assert np.array_equal(out, oracle.normalize_blocks(data, int(mode))), (n, mode)
            inplace = data.copy()
            ex.normalize_blocks(inplace, inplace, mode)   # in place is allowed (normalize.rs:44-50)
            assert np.array_equal(inplace, out)
        if n > 1:  # misaligned host buffers
            src = np.zeros(data.size + 1, np.uint8)
            src[1:] = data
            dst = np.zeros(data.size + 3, np.uint8)
            ex.normalize_blocks(src[1:], dst[3:], ex.ColorNormalizationMode.ReplicateColor)
            assert np.array_equal(dst[3:], oracle.normalize_blocks(data, 2))


def test_exhaustive_index_patterns_for_tricky_endpoints():
    """Every way of using the four index values (all 4^4 patterns over 4 pixel groups) x endpoint pairs where dictionary
    entries coincide or nearly coincide — the cases the kernel's fast reject must get right."""
    pairs = [(0xF800, 0xF800), (0x0000, 0x0000), (0xFFFF, 0xFFFF), (0x8410, 0x8410), (0x8410, 0x8411), (0x8411, 0x8410),
             (0x0001, 0x0000), (0x0000, 0x0001), (0x0020, 0x0000), (0x0800, 0x0000), (0xF800, 0x001F), (0x001F, 0xF800),
             (0x1234, 0x1235), (0x1235, 0x1234), (0xFFFF, 0x0000), (0x0000, 0xFFFF), (0x7BEF, 0x7BCF), (0x7BCF, 0x7BEF)]
    blocks = []
    for c0, c1 in pairs:
        for pat in range(256):
            idx = 0
            for g in range(4):   # each 2-bit digit of `pat` fills four pixels
                idx |= (((pat >> (2 * g)) & 3) * 0x55) << (8 * g)
            blocks.append(c0.to_bytes(2, "little") + c1.to_bytes(2, "little") + idx.to_bytes(4, "little"))
    data = np.frombuffer(b"".join(blocks), np.uint8).copy()
    outs = [np.zeros_like(data) for _ in range(3)]
    ex.normalize_blocks_all_modes(data, outs)
    want, _ = oracle.normalize_blocks_all_modes(data)
    for m in range(3):
        assert np.array_equal(outs[m], want[m]), m


def test_normalize_split_blocks_in_place():
    data = crafted_blocks(40_001, seed=5)
    n = data.size // 8
    for mode in MODES:
        colors, indices = data.reshape(n, 8)[:, :4].copy().reshape(-1), data.reshape(n, 8)[:, 4:].copy().reshape(-1)
        wc, wi = colors.copy(), indices.copy()
        ex.normalize_split_blocks_in_place(colors, indices, n, mode)
        oracle.normalize_split_blocks_in_place(wc, wi, int(mode))
        assert np.array_equal(colors, wc) and np.array_equal(indices, wi), mode


@pytest.mark.parametrize("n", [1, 77, 2048, 5463, 70_001])
def test_transform_with_normalize_blocks_all_24_combinations(n):
    """transform.rs:65 — normalization fused into the transform kernel (aligned, odd-N and byte-granular paths)."""
    import torch

    data = crafted_blocks(n, seed=100 + n)
    for d in ex.Bc1TransformDetailsWithNormalization.all_combinations():
        want = oracle.transform_with_normalize_blocks(data, int(d.color_normalization_mode), int(d.decorrelation_mode), d.split_colour_endpoints)
        out = np.zeros_like(data)
        ex.transform_bc1_with_normalize_blocks(data, out, d)
        assert np.array_equal(out, want), (n, d)
        # the untransform is the ordinary one and yields the NORMALIZED blocks
        back = np.zeros_like(data)
        dlt.untransform_bc1_with_settings(out, back, d.untransform_settings())
        assert np.array_equal(back, oracle.normalize_blocks(data, int(d.color_normalization_mode))), (n, d)
    # device-resident entry point, also at a misaligned device address (byte-granular kernel)
    d = ex.Bc1TransformDetailsWithNormalization(ex.ColorNormalizationMode.Color0Only, dlt.YCoCgVariant.Variant3, True)
    want = oracle.transform_with_normalize_blocks(data, 1, 3, True)
    for off in (0, 1):
        buf = torch.zeros(data.size + 16, dtype=torch.uint8, device="cuda")
        buf[off:off + data.size] = torch.from_numpy(data).cuda()
        dout = torch.zeros(data.size + 16, dtype=torch.uint8, device="cuda")
        ex.transform_bc1_with_normalize_blocks_device(buf.data_ptr() + off, dout.data_ptr() + off, data.size, d,
                                                      torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        assert np.array_equal(dout[off:off + data.size].cpu().numpy(), want), off


@pytest.mark.parametrize("use_all", [False, True])
def test_auto_with_normalization_matches_oracle(use_all):
    cases = [crafted_blocks(6_000, seed=9), crafted_blocks(30_011, seed=10), synth.texture_blocks(1, 9_000, seed=2)]
    # a smooth texture with flat and transparent regions pasted in
    tex = synth.texture_blocks(1, 20_000, seed=4).reshape(-1, 8).copy()
    tex[1000:3000] = np.frombuffer(bytes([0x00, 0xF8, 0x00, 0xF8, 0, 0, 0, 0]), np.uint8)
    tex[7000:9000] = np.frombuffer(bytes([0x00, 0x80, 0x00, 0xF8, 0xFF, 0xFF, 0xFF, 0xFF]), np.uint8)
    cases.append(tex.reshape(-1))
    for data in cases:
        out = np.zeros_like(data)
        best = ex.transform_bc1_auto_with_normalization(data, out, use_all)
        want_out, (nm, v, s) = oracle.auto_with_normalization(data, use_all)
        assert (int(best.color_normalization_mode), int(best.decorrelation_mode), best.split_colour_endpoints) == (nm, v, s)
        assert np.array_equal(out, want_out)
    # nothing normalizable: identical to the plain search
    rng = np.random.default_rng(3)
    plain = rng.integers(0, 256, 8 * 4000, dtype=np.uint8)
    _outs, any_ = oracle.normalize_blocks_all_modes(plain)
    if not any_:
        out, out2 = np.zeros_like(plain), np.zeros_like(plain)
        best = ex.transform_bc1_auto_with_normalization(plain, out, use_all)
        ref = dlt.transform_bc1_auto(plain, out2, dlt.Bc1EstimateSettings(dlt.LosslessTransformUtilsSizeEstimation(), use_all))
        assert best.color_normalization_mode == 0 and best.untransform_settings() == ref and np.array_equal(out, out2)


def test_one_gib_fused_normalization_keeps_the_roofline_shape():
    """Size-independent property at BASELINE scale: transform(normalize) == fused kernel, on 1 GiB, and both run."""
    import torch

    n = (1 << 30) // 8
    tile = torch.from_numpy(crafted_blocks(1 << 20, seed=77)).cuda()
    d_in = tile.repeat(n // (1 << 20))
    d_norm, d_a, d_b = torch.empty_like(d_in), torch.empty_like(d_in), torch.empty_like(d_in)
    stream = torch.cuda.current_stream().cuda_stream
    details = ex.Bc1TransformDetailsWithNormalization(ex.ColorNormalizationMode.ReplicateColor, dlt.YCoCgVariant.Variant1, True)
    ex.normalize_blocks_device(d_in.data_ptr(), d_norm.data_ptr(), d_in.numel(), details.color_normalization_mode, stream)
    dlt.transform_device(1, d_norm.data_ptr(), d_a.data_ptr(), d_in.numel(), details.untransform_settings(), stream)
    ex.transform_bc1_with_normalize_blocks_device(d_in.data_ptr(), d_b.data_ptr(), d_in.numel(), details, stream)
    torch.cuda.synchronize()
    assert torch.equal(d_a, d_b)
    assert not torch.equal(d_norm, d_in)
