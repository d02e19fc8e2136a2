#!/usr/bin/env python3
"""Build tuning variants of libmp3b200 (same sources, different -D knobs) for A/B timing in one GPU session.
usage: build_variants.py name=DEF1,DEF2 ...   ->  lamejs_b200/libmp3b200_<name>.so"""
import os
import sys
from concurrent.futures import ThreadPoolExecutor

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lamejs_b200 import build as B  # noqa: E402


def one(arg):
    name, _, defs = arg.partition("=")
    return B.build(variant=name, defines=[d for d in defs.split(",") if d])


with ThreadPoolExecutor(4) as ex:
    for p in ex.map(one, sys.argv[1:]):
        print(p)
