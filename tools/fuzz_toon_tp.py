"""Differential fuzz of the token-parallel TOON kernel body (csrc/json_tp.h, run on the CPU through the TEST-ONLY warp
emulator) against the sequential encoder (csrc/json_toon.h, itself pinned to the reference's golden vectors).
usage: python tools/fuzz_toon_tp.py [seed] [cases]"""
import json
import os
import random
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import hostsim_util as hs  # noqa: E402

KEYS = ["id", "name", "a", "b", "c", "x", "y", "key one", "k-2", "null", "true", "_p", "A.b", "é", "0k", "", "note", "value", "items"]
STRS = ["", "x", "hello world", "null", "true", "false", "12", "05", "1e5", "-a", "-", " lead", "trail ", "a,b", "a:b", "[x]", "{y}", "q\"uote",
        "back\\slash", "nl\nnl", "tab\t", "é", "日本語", "\U0001F600", "1٢", " x", "x ", "a/b", "ctrl\x01", "bell\b", "ff\f",
        "user1@example.com", "lorem ipsum dolor sit amet consectetur adipiscing elit sed do", "0", "-0", "1.0", "3.14", "0x1", "+1", ".5", "1.", "a" * 70,
        "é" * 40, "x-" * 40, "line1\nline2\nline3 with more text to make it long enough for the long span path \\ and a quote \" here"]


def make_gen(rng):
    def rstr():
        if rng.random() < 0.7:
            return rng.choice(STRS)
        return "".join(rng.choice("abc XYZ019_-.,:é\"\\\n") for _ in range(rng.randint(0, 12)))

    def rnum():
        r = rng.random()
        if r < 0.4:
            return rng.randint(-1000, 100000)
        if r < 0.6:
            return round(rng.uniform(-100, 100), rng.randint(0, 4))
        if r < 0.7:
            return rng.choice([0, -0.0, 0.0, 1e16, 1.5e-7, 1e-20, 123456789012345678, 2 ** 64, 2 ** 63, -2 ** 63, 0.1 + 0.2, 1 / 3, 1e22, 100.0, 12.50])
        if r < 0.8:
            return rng.random() * 10 ** rng.randint(-8, 8)
        return rng.randint(0, 9)

    def rprim():
        r = rng.random()
        if r < 0.4:
            return rstr()
        if r < 0.75:
            return rnum()
        return rng.choice([None, True, False])

    def rval(d):
        r = rng.random()
        if d <= 0 or r < 0.3:
            return rprim()
        if r < 0.45:
            return [rprim() for _ in range(rng.randint(0, 5))]
        if r < 0.6:  # table-like
            keys = rng.sample(KEYS, rng.randint(1, 4))
            rows = []
            for _ in range(rng.randint(1, 40 if rng.random() < 0.1 else 5)):
                ks = list(keys)
                q = rng.random()
                if q < 0.05:
                    rng.shuffle(ks)
                elif q < 0.09:
                    ks = ks[:-1] or ks
                elif q < 0.13:
                    ks = ks + [rng.choice(KEYS)]
                rows.append({k: (rprim() if rng.random() < 0.95 else rval(d - 1)) for k in ks})
            if rng.random() < 0.1:
                rows.insert(rng.randint(0, len(rows)), rprim())
            return rows
        if r < 0.8:
            return {rng.choice(KEYS): rval(d - 1) for _ in range(rng.randint(0, 5))}
        return [rval(d - 1) for _ in range(rng.randint(0, 4))]

    def dumps(v):
        r = rng.random()
        if r < 0.5:
            return json.dumps(v, separators=(",", ":"), ensure_ascii=False)
        if r < 0.7:
            return json.dumps(v, ensure_ascii=True)
        if r < 0.85:
            return json.dumps(v, indent=rng.choice([1, 2, 4]), ensure_ascii=False)
        return json.dumps(v, separators=(" , ", " : "), ensure_ascii=False)

    def mutate(t):
        if rng.random() < 0.75 or not t:
            return t
        b = bytearray(t.encode("utf-8", "surrogatepass"))
        k = rng.random()
        i = rng.randrange(len(b))
        if k < 0.3:
            del b[i]
        elif k < 0.6:
            b[i] = rng.choice(b',:{}[]"\\ 0-etn\x01\xe9')
        elif k < 0.8:
            b.insert(i, rng.choice(b',:{}[]"\\ 0-e'))
        else:
            b = b[:i]
        return bytes(b).decode("utf-8", "surrogateescape").encode("utf-8", "surrogatepass").decode("utf-8", "surrogatepass")

    def case():
        return mutate(dumps(rval(rng.randint(0, 5))))

    return case


def check(t, unlimited, rep, order):
    """None when the token-parallel result equals the sequential one (or the unit is handed to it), else a message."""
    a = hs.toon_host(t, unlimited=unlimited)
    b = hs.toon_tp(t, unlimited=unlimited, report_errors=rep, order=order)
    if b[0] == 7:
        return "fallback"
    if b[0] < 0:
        return f"NON-UNIFORM status {b}"
    if a == b:
        return None
    if not rep and a[0] in (1, 3, 4) and b[0] in (1, 3, 4):   # error vs not-smaller found in another order: the item is kept either way
        return None
    return f"seq={a} tp={b}"


def main():
    seed = int(sys.argv[1]) if len(sys.argv) > 1 else 1
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 2000
    rng = random.Random(seed)
    case = make_gen(rng)
    t0 = time.time()
    nfb = nbad = 0
    for it in range(n):
        t = case()
        r = check(t, rng.random() < 0.5, rng.random() < 0.5, (it & 1) | (rng.randrange(16) << 4))
        if r == "fallback":
            nfb += 1
        elif r:
            nbad += 1
            if nbad <= 8:
                print("BAD", repr(t)[:500], "\n   ", r)
    print(f"seed={seed} cases={n} bad={nbad} fallback={nfb} time={time.time() - t0:.1f}s")
    return 1 if nbad else 0


if __name__ == "__main__":
    sys.exit(main())
