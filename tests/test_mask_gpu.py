"""GPU parity of the request_logging_masking drop-in module (mask_kernel / classify_keys_kernel through
the C ABI) against the oracle, the crate's own unit tests and the vectors of the reference's Python twin."""
import importlib
import json
import os

import pytest

from mcp_context_forge_b200 import synth
from oracle import mask_ref

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden", "masking_twin.json")


@pytest.fixture(scope="module")
def mod():
    return importlib.import_module("request_logging_masking_native_extension")   # the name the middleware imports


@pytest.fixture(scope="module")
def gold():
    with open(GOLD, encoding="utf-8") as f:
        return json.load(f)


def test_crate_unit_tests_and_benchmark_parity_vectors(mod):
    assert mod.mask_sensitive_data({"password": "secret", "nested": {"authToken": "abc", "count": 3}}, 10) == {"password": "******", "nested": {"authToken": "******", "count": 3}}
    # tests/performance/test_request_logging_masking_native_extension_benchmark.py:174-190
    for v in [{"password": "secret", "nested": {"authToken": "abc", "ok": "value"}}, {"token_count": 3, "tokenizer": "ok", "privateKey": "secret"}, [{"jwt_token": "abc"}, {"normal": "value"}]]:
        assert mod.mask_sensitive_data(v, 12) == mask_ref.mask_value(v, 12)
    for h in [{"Authorization": "Bearer abc", "Cookie": "jwt_token=abc; theme=dark", "X-Trace-Id": "123"}, {"X-Auth-Count": "5", "X-Api-Key": "secret"}]:
        assert mod.mask_sensitive_headers(h) == mask_ref.mask_headers(h)
    assert mod.mask_sensitive_headers({"Cookie": "jwt_token=abc; theme=dark; session_id=xyz"}) == {"Cookie": "jwt_token=******; theme=dark; session_id=******"}
    with pytest.raises(TypeError):
        mod.mask_sensitive_headers(["x"])
    with pytest.raises(ValueError):
        mod.mask_sensitive_json_bytes(b'{"a":')
    assert mod.mask_sensitive_json_bytes(b'{"level":{"nested":{}}}', 1) == b'{"level":"<nested too deep>"}'


def test_twin_golden_object_api(mod, gold):
    for c in gold["mask_sensitive_data"]:
        assert mod.mask_sensitive_data(c["data"], c["max_depth"]) == c["masked"]
    for c in gold["headers"]:
        if all(isinstance(v, str) and v.isascii() and "\x1c" not in v for v in c["headers"].values()) and all(k.isascii() for k in c["headers"]):
            assert mod.mask_sensitive_headers(c["headers"]) == c["masked"]


def test_bytes_batch_vs_oracle(mod, gold):
    payloads = [json.dumps(c["data"]).encode() for c in gold["mask_sensitive_data"][::4]]
    payloads += [synth.payload("B", 16384, seed=s).encode() for s in range(64)] + [synth.payload("A", 16384, seed=s).encode() for s in range(32)]
    payloads += [synth.payload("B", 262144, seed=s).encode() for s in range(8)] + [b'{"a":1,}', b"[01]", b"\xff", b"[1e400]", b"", b'[0.30000000000000004,5e-324,1e23,-0]']
    for md in (10, 3):
        got = mod.mask_sensitive_json_bytes_batch(payloads, md)
        for p, g in zip(payloads, got):
            try:
                exp = mask_ref.mask_json_bytes(p, md)
            except ValueError:
                exp = None
            assert g == exp, p[:120]


def test_benchmark_scenario_payload(mod):
    """The reference benchmark's nested payload (…benchmark.py:63-79), 1024 events, through the bytes API."""
    payload = {"events": [{"actor": {"userName": f"user-{i}", "sessionToken": f"token-{i}", "sessionCount": i},
                           "request": {"clientSecret": f"secret-{i}", "payload": {"safeField": "value" * 8, "authDevice": f"device-{i}", "auth_count": i}}} for i in range(1024)]}
    raw = json.dumps(payload).encode()
    assert mod.mask_sensitive_json_bytes(raw, 12) == mask_ref.mask_json_bytes(raw, 12)
    assert mod.mask_sensitive_data(payload, 12) == mask_ref.mask_value(payload, 12)


def test_non_json_fallback_and_header_batch():
    """the middleware's non-JSON branch (13 lowercase substring probes, request_logging_middleware.py:661-667) and
    the header path for a wave of requests, against the reference's own expressions evaluated in Python."""
    import random

    from mcp_context_forge_b200 import masking

    rng = random.Random(4)
    words = ["hello", "TOKEN", "Api_Key", "user=bob", "pass word", "PassPhrase", "refresh_token=1", "tokKen", "toKen", "apİ_key", "secreté", "AUTHORIZATION:",
             "client-secret", "private_key", "plain text", "x" * 200, "jwt_TOKEN", "Auth_Token"]
    bodies = []
    for _ in range(300):
        b = " ".join(rng.choice(words) for _ in range(rng.randint(0, 6))).encode("utf-8")
        if rng.random() < 0.2:
            i = rng.randrange(len(b) + 1)
            b = b[:i] + rng.choice([b"\xff", b"\xc3", b"\xe2\x84"]) + b[i:]           # invalid UTF-8: errors="ignore" drops it
        bodies.append(b)
    bodies += [b"", b"tok\xffen", b"TOK\xc4\xb0EN"]
    got = masking.non_json_fallback_batch(bodies)
    for b, g in zip(bodies, got):
        s = b.decode("utf-8", errors="ignore")
        exp = "<contains sensitive data - masked>" if any(k in s.lower() for k in masking.SENSITIVE_KEYS) else s
        assert g == exp, (b, g, exp)
    assert sum(1 for g in got if g == masking.NON_JSON_MASKED) > 50
    hs = [{"Content-Type": "json", "Authorization": "Bearer x", "Cookie": "a=1; session_id=9", "X-Api-Key": "k", "X-Token-Count": "3"}, {"cookie": "jwt=1", "Accept": "*/*"}, {}]
    assert masking.mask_sensitive_headers_batch(hs) == [masking.mask_sensitive_headers(h) for h in hs]
