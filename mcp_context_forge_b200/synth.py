"""Deterministic synthetic tool-call payloads: shape A "tabular", shape B "nested
config", shape C "prose" with injected pattern hits.  Used by bench.py and the tests; no I/O."""
from __future__ import annotations

import json
import random
from typing import List

LOREM = "lorem ipsum dolor sit amet consectetur adipiscing elit sed do"
HARMFUL_PHRASES = ["kill myself", "suicide", "self-harm", "want to die", "kill him", "shoot them", "stab someone",
                   "eradicate those people", "hate speech", "racial slur", "Kill her", "ſuicide"]
DENY_WORDS = ["innovative", "groundbreaking", "revolutionary"]
REGEX_WORDS = ["crap", "crud"]
NON_ASCII = ["é", "ß", "ſ", "K", "日本語", "😀", "naïve", "Ünïcödé"]


def _vocab(rng: random.Random, n: int = 5000) -> List[str]:
    letters = "abcdefghijklmnopqrstuvwxyz"
    return ["".join(rng.choice(letters) for _ in range(rng.randint(2, 10))) for _ in range(n)]


def shape_a(target_bytes: int, seed: int = 0) -> str:
    """{"results":[{id,name,email,score,active,note}...],"meta":{...}} compact JSON >= target bytes."""
    rng = random.Random(seed)
    rows = []
    size = 40
    i = 0
    while size < target_bytes:
        row = {"id": i, "name": f"user{i}", "email": f"user{i}@example.com", "score": round(rng.uniform(0, 100), 2),
               "active": bool(rng.getrandbits(1)), "note": LOREM[: 54]}
        rows.append(row)
        size += len(json.dumps(row, separators=(",", ":"))) + 1
        i += 1
    return json.dumps({"results": rows, "meta": {"count": len(rows), "source": "db"}}, separators=(",", ":"))


def shape_b(target_bytes: int, seed: int = 0) -> str:
    """Nested config-like JSON, ~10 % sensitive keys, ~10 % keys with non-sensitive suffixes."""
    rng = random.Random(seed)
    sens = ["password", "authToken", "client_secret", "X-Api-Key", "sessionToken"]
    nons = ["token_count", "auth_status", "secret_name", "password_length"]
    plain = ["name", "host", "port", "enabled", "retries", "path", "mode", "level", "tags", "items", "opts", "limits"]

    def node(depth: int):
        r = rng.random()
        if depth <= 0 or r < 0.35:
            t = rng.random()
            if t < 0.4:
                return "".join(rng.choice("abcdefghij") for _ in range(rng.randint(3, 12)))
            if t < 0.7:
                return rng.randint(0, 100000)
            if t < 0.8:
                return round(rng.uniform(-100, 100), 3)
            if t < 0.9:
                return bool(rng.getrandbits(1))
            return None
        if r < 0.55:
            return [node(depth - 1) for _ in range(rng.randint(1, 4))]
        d = {}
        for _ in range(rng.randint(2, 6)):
            kr = rng.random()
            k = rng.choice(sens) if kr < 0.1 else rng.choice(nons) if kr < 0.2 else rng.choice(plain) + str(rng.randint(0, 9))
            d[k] = node(depth - 1)
        return d

    out = {}
    size = 2
    i = 0
    while size < target_bytes:
        v = node(rng.randint(4, 8))
        out[f"section{i}"] = v
        size += len(json.dumps(v, separators=(",", ":"))) + 12
        i += 1
    return json.dumps(out, separators=(",", ":"))


def shape_c(target_bytes: int, seed: int = 0, hit_rate: float = 1e-4, non_ascii_rate: float = 0.01) -> str:
    """English-like prose (Zipf over a 5k vocabulary) with injected hits and non-ASCII words."""
    rng = random.Random(seed)
    vocab = _vocab(random.Random(12345))
    weights = [1.0 / (i + 1) for i in range(len(vocab))]
    inject = HARMFUL_PHRASES + DENY_WORDS + REGEX_WORDS
    words: List[str] = []
    size = 0
    chunk = rng.choices(vocab, weights, k=4096)
    ci = 0
    while size < target_bytes:
        r = rng.random()
        if r < hit_rate:
            w = rng.choice(inject)
        elif r < hit_rate + non_ascii_rate:
            w = rng.choice(NON_ASCII)
        else:
            if ci == len(chunk):
                chunk = rng.choices(vocab, weights, k=4096)
                ci = 0
            w = chunk[ci]
            ci += 1
        words.append(w)
        size += len(w.encode()) + 1
    return " ".join(words)


def payload(shape: str, target_bytes: int, seed: int = 0, hit_rate: float = 1e-4) -> str:
    if shape == "A":
        return shape_a(target_bytes, seed)
    if shape == "B":
        return shape_b(target_bytes, seed)
    if shape == "C":
        return shape_c(target_bytes, seed, hit_rate)
    raise ValueError(shape)
