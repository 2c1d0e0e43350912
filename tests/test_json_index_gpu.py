"""GPU parity of the JSON structural index kernel (cf_json_index_host through the C ABI) against the host
build of the same header (tests/hostsim), which tests/test_json_index_cpu.py pins to an independent Python
restatement and to the sequential parser.  Bit-exact token lists, with and without classification."""
import numpy as np
import pytest

import hostsim_util as hs
from mcp_context_forge_b200 import engine, synth
from test_json_index_cpu import EDGE, corpus, py_tokens

pytestmark = pytest.mark.gpu


def run(units, classify):
    ctx = engine.Context.get()
    stream, offs = engine.pack_units(units)
    batch = engine.Batch(ctx, len(stream), len(units))
    return engine.json_index_host(batch, stream, offs, classify)


@pytest.mark.parametrize("classify", [False, True])
def test_index_matches_host_build(classify):
    units = [e for e in EDGE if b"\xff" not in e] + [t.encode("utf-8", "surrogatepass") for t in corpus()]
    got = run(units, classify)
    assert len(got) == len(units)
    for data, (toks, unt) in zip(units, got):
        exp, exp_unt = hs.json_index(data)
        assert unt == exp_unt, data[:100]
        assert [(int(p) & 0x7FFFFFFF, bool(int(p) >> 31)) for p, _ in toks] == [(p, c) for p, c, _ in exp], data[:100]
        # ... and DIRECTLY against the independent Python restatement of the token definition (not only the shared header)
        ptoks, punt = py_tokens(data)
        assert [(int(p) & 0x7FFFFFFF, bool(int(p) >> 31)) for p, _ in toks] == list(ptoks) and unt == punt, data[:100]
        if classify:
            assert [int(a) for _, a in toks] == [a for _, _, a in exp], data[:100]


def test_index_full_size_units_and_chunk_boundaries():
    """16 KiB and 256 KiB units, plus strings/escapes straddling every 32-byte chunk boundary."""
    units = [synth.payload(s, size, seed=k).encode() for s in "ABC" for size, k in ((16384, 1), (262144, 2))]
    for pad in range(0, 70):
        units.append(b" " * pad + b'["a\\\\\\"b\\\\", "' + b"\\\\" * 17 + b'", {"k": -1.5e3}]')
    got = run(units, True)
    for data, (toks, unt) in zip(units, got):
        exp, exp_unt = hs.json_index(data)
        assert unt == exp_unt
        assert [(int(p), int(a)) for p, a in toks] == [(p | (0x80000000 if c else 0), a) for p, c, a in exp]
