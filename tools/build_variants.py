"""Build experiment variants of libfa_b200.so next to the default one: python tools/build_variants.py name=FLAG,FLAG ...
e.g.  python tools/build_variants.py r1=-DFA_FWD_EXP_PIPE=0 p3=-DFA_FWD_POLY8=3   ->  flash_attention_b200/variants/lib_<name>.so"""
import concurrent.futures as cf
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from flash_attention_b200 import build as b  # noqa: E402

out_dir = os.path.join(b.HERE, "variants")
os.makedirs(out_dir, exist_ok=True)


def one(spec):
    name, _, flags = spec.partition("=")
    out = os.path.join(out_dir, f"lib_{name}.so")
    b.build(force=True, extra_flags=tuple(f for f in flags.split(",") if f), out=out)
    return out


with cf.ThreadPoolExecutor(4) as ex:
    for p in ex.map(one, sys.argv[1:]):
        print("built", p, flush=True)
