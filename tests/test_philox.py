"""The render's random stream: Philox4x32 (Salmon et al., SC11). The oracle's implementation is pinned to the published
known-answer vectors of the Random123 distribution (kat_vectors); the device implementation (expanded key schedule, 23-bit
uniforms from a mantissa trick) must produce the oracle's blocks and uniforms bit for bit."""
import numpy as np
import pytest

from ipt_b200 import capi

# Random123 kat_vectors: philox4x32 <rounds> <counter x4> <key x2> -> <output x4>
KAT = [
    (10, [0, 0, 0, 0], [0, 0], [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]),
    (10, [0xFFFFFFFF] * 4, [0xFFFFFFFF] * 2, [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]),
    (10, [0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344], [0xA4093822, 0x299F31D0], [0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1]),
    (7, [0, 0, 0, 0], [0, 0], [0x5F6FB709, 0x0D893F64, 0x4F121F81, 0x4F730A48]),
]


@pytest.mark.parametrize("rounds,counter,key,expected", KAT)
def test_oracle_philox_known_answers(rounds, counter, key, expected, oracle):
    assert oracle.philox_rounds(rounds, counter, key) == expected


def _counters(n, seed):
    rng = np.random.default_rng(seed)
    c = rng.integers(0, 2**32, size=(n, 4), dtype=np.uint64).astype(np.uint32)
    c[:8] = [[0, 0, 0, 0], [0xFFFFFFFF] * 4, [1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 1, 0], [0, 0, 0, 1], [409599, 1023, 2047, 3], [5, 6, 7, 8]]
    return c


@pytest.mark.gpu
@pytest.mark.parametrize("seed", [0, 2024, 0xFFFFFFFFFFFFFFFF, 0x0123456789ABCDEF])
def test_device_philox_equals_oracle(seed, lib, oracle):
    c = _counters(4096, 11)
    blocks, uni = capi.philox_batch(c, seed)
    k = [seed & 0xFFFFFFFF, seed >> 32]
    ref = np.array([oracle.philox(list(map(int, row)), k) for row in c], np.uint32)
    assert np.array_equal(blocks, ref)
    # uniforms: (x >> (32 - bits)) * 2^-bits, never 1.0 (include/randf.h:6-11)
    bits = 23
    expect = (ref >> np.uint32(32 - bits)).astype(np.float32) * np.float32(1.0 / (1 << bits))
    assert np.array_equal(uni.view(np.uint32), expect.view(np.uint32))
    assert uni.max() < 1.0 and uni.min() >= 0.0
