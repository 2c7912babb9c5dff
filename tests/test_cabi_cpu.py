"""CPU-side checks: the C-ABI library builds/loads without a GPU and exports every symbol that
include/fa_b200.h declares; host-only helpers behave like the reference's."""
import ctypes as C
import os
import re

from flash_attention_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "fa_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(fa_b200_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol(fa_lib):
    names = _declared_symbols()
    assert len(names) >= 12
    for n in names:
        assert hasattr(fa_lib, n), n
        assert n in _lib.SYMBOLS, f"{n} missing from the ctypes binding table"


def test_struct_abi_sizes_match(fa_lib):
    assert fa_lib.fa_b200_sizeof_fwd_args() == C.sizeof(_lib.FwdArgs)
    assert fa_lib.fa_b200_sizeof_bwd_args() == C.sizeof(_lib.BwdArgs)


def test_version_and_launch_counter(fa_lib):
    assert b"sm_100a" in fa_lib.fa_b200_version()
    assert fa_lib.fa_b200_launch_count() >= 0


def test_num_splits_heuristic_matches_reference_rule(fa_lib):
    """flash_api.cpp:281-315: 1 split when the grid already fills >= 80 % of the SMs; otherwise the
    smallest split count within 85 % of the best wave efficiency; never more than the block count."""
    f = fa_lib.fa_b200_num_splits_heuristic
    assert f(148, 148, 64, 128) == 1
    assert f(119, 148, 64, 128) == 1  # 0.8 * 148 = 118.4
    assert f(1, 148, 1, 128) == 1
    assert f(1, 148, 4, 128) == 4
    s = f(8, 148, 512, 128)
    assert 1 < s <= 128 and (8 * s / 148) / -(-8 * s // 148) >= 0.85 * 0.999
    # python restatement of the rule, all small cases
    import math
    def ref(ctas, sms, nb, mx):
        if ctas >= 0.8 * sms: return 1
        mx = min(mx, sms, nb)
        if mx <= 1: return 1
        cd = lambda a, b: -(-a // b)
        elig = lambda s: s == 1 or cd(nb, s) != cd(nb, s - 1)
        eff = lambda s: (ctas * s / sms) / math.ceil(ctas * s / sms)
        best = max(eff(s) for s in range(1, mx + 1) if elig(s))
        for s in range(1, mx + 1):
            if elig(s) and eff(s) >= 0.85 * best: return s
        return 1
    for ctas in (1, 2, 3, 7, 16, 40, 64, 100):
        for nb in (1, 2, 5, 17, 64, 512):
            assert f(ctas, 148, nb, 128) == ref(ctas, 148, nb, 128), (ctas, nb)


def test_argument_errors_are_reported_without_gpu(fa_lib):
    a = _lib.FwdArgs()
    assert fa_lib.fa_b200_mha_fwd(C.byref(a), None) != 0
    assert b"fp16 and bf16" in fa_lib.fa_b200_last_error() or b"non-null" in fa_lib.fa_b200_last_error()
    a.dtype = _lib.BF16
    a.q = a.k = a.v = a.out = a.softmax_lse = 16
    a.b, a.seqlen_q, a.seqlen_k, a.h, a.h_k, a.d = 1, 128, 128, 6, 4, 128
    assert fa_lib.fa_b200_mha_fwd(C.byref(a), None) != 0
    assert b"must divide" in fa_lib.fa_b200_last_error()
    a.h_k = 2
    a.d = 100
    assert fa_lib.fa_b200_mha_fwd(C.byref(a), None) != 0
    assert b"multiple of 8" in fa_lib.fa_b200_last_error()
    assert fa_lib.fa_b200_mha_fwd(None, None) != 0
