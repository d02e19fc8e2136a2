"""Committed golden fixtures (tests/golden/golden.json, produced by the oracle): the oracle must keep reproducing
them on CPU; the GPU encoder must reproduce them through the C-ABI on the B200."""
import hashlib
import json
import os

import pytest

from synth import make_signal

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, "golden", "golden.json")))


@pytest.mark.parametrize("name", sorted(GOLD))
def test_oracle_reproduces_golden(oracle, name):
    g = GOLD[name]
    l, r = make_signal(g["kind"], g["samples"], g["samplerate"], g["seed"])
    data, _, _ = oracle.encode_stream(g["channels"], g["samplerate"], g["kbps"], l, r if g["channels"] == 2 else None)
    assert len(data) == g["bytes"] and hashlib.sha256(data).hexdigest() == g["sha256"]


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(GOLD))
def test_gpu_reproduces_golden(name):
    import lamejs_b200 as M

    g = GOLD[name]
    l, r = make_signal(g["kind"], g["samples"], g["samplerate"], g["seed"])
    out = M.encode_streams(g["channels"], g["samplerate"], g["kbps"], [l], [r])[0]
    assert len(out) == g["bytes"]
    assert out[:48].hex() == g["head"]
    assert hashlib.sha256(out).hexdigest() == g["sha256"]
