"""CPU tests of the oracle's restatement of experimental::normalize_blocks against the reference's own test vectors
(core/dxt-lossless-transform-bc1/src/experimental/normalize_blocks/normalize.rs tests) and its structural properties."""
import numpy as np
import pytest

import oracle
from norm_cases import REFERENCE_VECTORS, crafted_blocks


@pytest.mark.parametrize("mode", [1, 2])
def test_reference_vectors(mode):
    for name, block, expected in REFERENCE_VECTORS:
        src = np.frombuffer(block, np.uint8).copy()
        want = expected[mode] if expected[mode] is not None else block
        assert oracle.normalize_blocks(src, mode).tobytes() == want, name
        assert oracle.normalize_blocks(src, 0).tobytes() == block, name  # mode None copies


def test_multiple_blocks_and_all_modes():
    # normalize.rs `can_normalize_multiple_blocks` / `can_normalize_blocks_all_modes`: solid red then transparent
    src = np.frombuffer(REFERENCE_VECTORS[4][1] + REFERENCE_VECTORS[1][1], np.uint8).copy()
    outs, any_ = oracle.normalize_blocks_all_modes(src)
    assert any_
    assert outs[0].tobytes() == src.tobytes()[:8] + b"\xFF" * 8          # None: solid block copied, transparent block 0xFF
    assert outs[1].tobytes() == bytes([0x00, 0xF8]) + bytes(6) + b"\xFF" * 8
    assert outs[2].tobytes() == bytes([0x00, 0xF8, 0x00, 0xF8]) + bytes(4) + b"\xFF" * 8
    mixed = np.frombuffer(REFERENCE_VECTORS[2][1] * 3, np.uint8).copy()
    outs, any_ = oracle.normalize_blocks_all_modes(mixed)
    assert not any_ and all(o.tobytes() == mixed.tobytes() for o in outs)


def test_decoder_known_answers():
    # util/bc1_decode.rs tests: red endpoints decode to (255, 0, 0, 255); 0x8000 -> r = (16 << 3) | (16 >> 2) = 132
    blk = np.frombuffer(bytes([0x00, 0xF8, 0x00, 0xF8, 0, 0, 0, 0]), np.uint8).copy()
    assert oracle.normalize_blocks(blk, 1).tobytes() == bytes([0x00, 0xF8]) + bytes(6)      # solid, round-trips
    grey = np.frombuffer(bytes([0x10, 0x84, 0x10, 0x84, 0x55, 0x55, 0x55, 0x55]), np.uint8).copy()  # equal endpoints, index 1
    assert oracle.normalize_blocks(grey, 2).tobytes() == bytes([0x10, 0x84, 0x10, 0x84]) + bytes(4)
    # equal endpoints put the block in punch-through mode: index 3 is transparent -> all 0xFF
    tr = np.frombuffer(bytes([0x10, 0x84, 0x10, 0x84, 0xFF, 0xFF, 0xFF, 0xFF]), np.uint8).copy()
    assert oracle.normalize_blocks(tr, 1).tobytes() == b"\xFF" * 8


def test_structural_properties_on_crafted_blocks():
    data = crafted_blocks(50_000, seed=1)
    n = data.size // 8
    outs, any_ = oracle.normalize_blocks_all_modes(data)
    assert any_
    changed = [(outs[m].reshape(n, 8) != data.reshape(n, 8)).any(axis=1).sum() for m in range(3)]
    assert changed[1] > n // 10 and changed[2] > n // 10, changed     # the generator really produces normalizable blocks
    for m in range(3):
        assert np.array_equal(oracle.normalize_blocks(data, m) if m else outs[0] * 0 + oracle.normalize_blocks_all_modes(data)[0][0], outs[m])
    for m in (1, 2):
        # idempotent
        assert np.array_equal(oracle.normalize_blocks(outs[m], m), outs[m])
        # split-in-place == normalize on whole blocks
        blocks = data.reshape(n, 8)
        colors, indices = blocks[:, :4].copy().reshape(-1), blocks[:, 4:].copy().reshape(-1)
        oracle.normalize_split_blocks_in_place(colors, indices, m)
        want = outs[m].reshape(n, 8)
        assert np.array_equal(colors.reshape(n, 4), want[:, :4]) and np.array_equal(indices.reshape(n, 4), want[:, 4:])
    # transform_bc1_with_normalize_blocks == transform(normalize(x)) for every combination
    for m in range(3):
        for v in range(4):
            for s in (False, True):
                got = oracle.transform_with_normalize_blocks(data, m, v, s)
                assert np.array_equal(got, oracle.transform(1, outs[m] if m else data, v, False, s)), (m, v, s)


@pytest.mark.parametrize("use_all", [False, True])
def test_auto_with_normalization_oracle(use_all):
    # nothing normalizable -> the plain search
    rng = np.random.default_rng(3)
    plain = rng.integers(0, 256, 8 * 3000, dtype=np.uint8)
    plain.reshape(-1, 8)[:, 4:] = rng.integers(1, 255, (3000, 4), dtype=np.uint8)
    outs, any_ = oracle.normalize_blocks_all_modes(plain)
    if not any_:
        out, (nm, v, s) = oracle.auto_with_normalization(plain, use_all)
        want_out, (wv, _sa, ws) = oracle.auto(1, plain, use_all)
        assert nm == 0 and (v, s) == (wv, ws) and np.array_equal(out, want_out)
    # normalizable blocks -> 3 x K candidates, result is one of the fused transforms
    data = crafted_blocks(6000, seed=9)
    out, (nm, v, s) = oracle.auto_with_normalization(data, use_all)
    assert np.array_equal(out, oracle.transform_with_normalize_blocks(data, nm, v, s))
    back = oracle.untransform(1, out, v, False, s)
    assert np.array_equal(back, oracle.normalize_blocks(data, nm))
