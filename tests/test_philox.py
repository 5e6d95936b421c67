"""Known-answer vectors of Philox4x32-10 (Salmon et al., SC'11; Random123 kat_vectors)."""
import pytest

from oracle import oracle as orc
from oracle import ref_runner as rr

KAT = [
    ((0, 0), (0, 0, 0, 0), (0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8)),
    ((0xFFFFFFFF, 0xFFFFFFFF), (0xFFFFFFFF,) * 4, (0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD)),
    ((0xA4093822, 0x299F31D0), (0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344),
     (0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1)),
]


@pytest.mark.parametrize("key,ctr,want", KAT)
def test_oracle_c_philox(key, ctr, want):
    assert tuple(orc.philox(*key, *ctr)) == want


@pytest.mark.parametrize("key,ctr,want", KAT)
def test_python_tape_philox(key, ctr, want):
    assert tuple(rr.philox4x32_10(*key, *ctr)) == want
