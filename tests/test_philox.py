"""Philox4x32-10 known-answer vectors (Random123 kat_vectors) for the oracle copy and the device
copy, and agreement of the uniform draw schedule between the two."""
import numpy as np
import pytest

from oracle import orcapi

KAT = [
    ([0, 0, 0, 0], [0, 0], [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]),
    ([0xffffffff] * 4, [0xffffffff] * 2, [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]),
    ([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0],
     [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]),
]


def test_oracle_philox_kat():
    for ctr, key, want in KAT:
        assert orcapi.philox(ctr, key) == want


def test_oracle_uniform_range_and_keys():
    draws = [orcapi.uniform(1234, c, it, l, s) for c in range(2) for it in range(3) for l in range(3) for s in range(5)]
    assert all(0.0 <= d < 1.0 for d in draws)
    assert len(set(draws)) == len(draws)
    assert abs(np.mean([orcapi.uniform(9, 0, i, 0, 0) for i in range(4000)]) - 0.5) < 0.02


@pytest.mark.gpu
def test_device_philox_kat():
    from swiftlink_b200 import capi
    for ctr, key, want in KAT:
        assert capi.debug_philox(ctr, key) == want


@pytest.mark.gpu
def test_device_uniform_matches_oracle():
    from swiftlink_b200 import capi
    for args in [(1, 0, 0, 0, 0), (2 ** 40 + 17, 3, 2 ** 33 + 5, 9999, 401), (77, 1, 12, 5, 0x7ffffff0), (5, 2, 7, 3, 8)]:
        assert capi.debug_uniform(*args) == orcapi.uniform(*args)
