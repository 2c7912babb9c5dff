"""CPU checks of the dropout RNG restatement (oracle/attention_oracle.py: philox4x32_7, philox_keep_mask).

The generator is pinned to the published Random123 known-answer vectors for Philox4x32 (Salmon et al., SC'11,
kat_vectors: the 10-round and the 7-round entries); the reference's csrc/flash_attn/src/philox.cuh:29-51 is
the same round function run 7 times."""
import numpy as np
import torch

from oracle.attention_oracle import attention_oracle, philox4x32_7, philox_keep_mask

KAT = [  # rounds, counter, key, expected
    (10, (0, 0, 0, 0), (0, 0), (0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8)),
    (10, (0xFFFFFFFF,) * 4, (0xFFFFFFFF,) * 2, (0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD)),
    (10, (0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344), (0xA4093822, 0x299F31D0),
     (0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1)),
    (7, (0, 0, 0, 0), (0, 0), (0x5F6FB709, 0x0D893F64, 0x4F121F81, 0x4F730A48)),
]


def test_philox_known_answers():
    for rounds, ctr, key, want in KAT:
        got = philox4x32_7(*[[c] for c in ctr], *key, rounds=rounds)
        assert tuple(int(x[0]) for x in got) == want


def test_keep_mask_shape_rate_and_determinism():
    m = philox_keep_mask(1234, 64, 0.17, 2, 3, 130, 70)
    assert m.shape == (2, 3, 130, 70) and m.dtype == torch.bool
    # 8-bit threshold: P(keep) = (floor(0.83 * 255) + 1) / 256
    assert abs(m.float().mean().item() - (int(np.floor(np.float32(0.83) * 255)) + 1) / 256) < 0.01
    assert torch.equal(m, philox_keep_mask(1234, 64, 0.17, 2, 3, 130, 70))
    assert not torch.equal(m, philox_keep_mask(1234, 68, 0.17, 2, 3, 130, 70))
    assert not torch.equal(m, philox_keep_mask(1235, 64, 0.17, 2, 3, 130, 70))
    # the mask of a sub-problem is the corner of the larger one (pure function of the coordinates)
    assert torch.equal(m[:, :, :50, :33], philox_keep_mask(1234, 64, 0.17, 2, 3, 50, 33))


def test_oracle_dropout_matches_manual():
    torch.manual_seed(0)
    q, k, v = torch.randn(1, 8, 2, 16), torch.randn(1, 8, 2, 16), torch.randn(1, 8, 2, 16)
    keep = philox_keep_mask(7, 0, 0.25, 1, 2, 8, 8)
    out, _ = attention_oracle(q, k, v, dropout_p=0.25, dropout_mask=keep)
    p = torch.softmax(torch.einsum("bihd,bjhd->bhij", q, k) / 4.0, dim=-1)
    want = torch.einsum("bhij,bjhd->bihd", p * keep / 0.75, v)
    assert torch.allclose(out, want, atol=1e-5)
