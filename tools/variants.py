#!/usr/bin/env python
"""Build experimental variants of libuvt.so (compile-time knobs) next to the product library, for A/B runs on the GPU box.

    python tools/variants.py name1:DEF1=V,DEF2=V name2:DEF=V ...      -> variants/libuvt_<name>.so
    (on the GPU box)  tools/bench3.sh <tag> UVT_LIB_PATH=variants/libuvt_<name>.so
"""
import importlib
import os
import sys
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
b = importlib.import_module("unnamed-voxel-tracer_b200.build")
os.makedirs(os.path.join(ROOT, "variants"), exist_ok=True)


def one(spec):
    name, _, defs = spec.partition(":")
    out = os.path.join(ROOT, "variants", f"libuvt_{name}.so")
    b.build(out=out, defines=[d for d in defs.split(",") if d])
    return out


with ThreadPoolExecutor(max_workers=4) as ex:
    for o in ex.map(one, sys.argv[1:]):
        print(o)
