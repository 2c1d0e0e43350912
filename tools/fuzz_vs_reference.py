"""Container-only: live differential fuzz of the TOON kernels' source (sequential encoder csrc/json_toon.h on the host build; token-parallel kernel
body csrc/json_tp.h on the 32-fibre warp emulator) against the REFERENCE'S OWN `plugins/toon_encoder/toon.py`, imported unmodified from
/root/reference — no restatement in between (the oracle is compared too, so a gap in it shows).  Random JSON documents from the generator of
tools/fuzz_toon_tp.py (adversarial keys / strings / numbers, tables, byte-level mutations).  orjson is not installable here: the strict stdlib
parser stands in, and documents that would expose an orjson / json delta (integers beyond 64 bits, lone surrogates, non-finite floats) are skipped.
usage: python tools/fuzz_vs_reference.py [seed] [cases]"""
import importlib.util
import json
import math
import os
import random
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import hostsim_util as hs  # noqa: E402
from fuzz_toon_tp import make_gen  # noqa: E402
from mcp_context_forge_b200.plugins.toon_encoder import _encode_error  # noqa: E402
from oracle import toon_ref  # noqa: E402


def load_reference_toon():
    spec = importlib.util.spec_from_file_location("ref_toon", os.path.join(REF, "plugins/toon_encoder/toon.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def delta_free(v) -> bool:
    """True when stdlib json and orjson agree on this parsed document."""
    if isinstance(v, bool) or v is None:
        return True
    if isinstance(v, int):
        return -(2 ** 63) <= v < 2 ** 64
    if isinstance(v, float):
        return math.isfinite(v)
    if isinstance(v, str):
        return not any(0xD800 <= ord(c) <= 0xDFFF for c in v)
    if isinstance(v, list):
        return all(delta_free(x) for x in v)
    return all(delta_free(k) and delta_free(x) for k, x in v.items())


def expected(toon, t: str):
    """(status, text) as include/cfgpu.h defines them for an unlimited output buffer: 0 = the reference's toon.encode(orjson.loads(t)),
    2 = not JSON, 3 / 4 = toon.encode raises ValueError / AttributeError."""
    try:
        t.encode("utf-8")
    except UnicodeEncodeError:
        return None
    try:
        doc = json.loads(t, parse_constant=lambda c: (_ for _ in ()).throw(ValueError(c)))
    except (ValueError, RecursionError):
        return (2, None)
    if not delta_free(doc):
        return None
    try:
        return (0, toon.encode(doc))
    except ValueError as exc:
        expected.message = str(exc)
        return (3, None)
    except AttributeError as exc:
        expected.message = str(exc)
        return (4, None)


def main() -> int:
    if not os.path.isdir(REF):
        print("fuzz_vs_reference: /root/reference is not here (container-only tool)")
        return 0
    seed = int(sys.argv[1]) if len(sys.argv) > 1 else 1
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 2000
    toon = load_reference_toon()
    rng = random.Random(seed)
    case = make_gen(rng)
    t0 = time.time()
    done = skipped = handed = bad = worded = 0
    for it in range(n):
        t = case()
        exp = expected(toon, t)
        if exp is None or (exp[0] == 2 and any(0xD800 <= ord(c) <= 0xDFFF for c in t)):
            skipped += 1
            continue
        done += 1
        seq = hs.toon_host(t, unlimited=True)
        tp = hs.toon_tp(t, unlimited=True, report_errors=True, order=(it & 1) | (rng.randrange(16) << 4))
        if tp[0] == 7:
            handed += 1
            tp = exp
        try:
            orc = (0, toon_ref.encode(toon_ref.loads_strict(t)))
        except ValueError:
            orc = (2 if exp[0] == 2 else 3, None)
        except toon_ref.ToonCrash:
            orc = (4, None)
        if exp[0] in (3, 4):                                    # the wording the drop-in gives the exception (skip_on_error: false) == the reference's
            worded += 1
            msg = str(_encode_error(exp[0], t))
            if msg != expected.message:
                bad += 1
                if bad <= 8:
                    print("BAD MESSAGE", repr(t)[:400], "\n   reference", expected.message, "\n   drop-in  ", msg)
        if not (seq == exp and tp == exp and orc == exp):
            bad += 1
            if bad <= 8:
                print("BAD", repr(t)[:400], "\n   reference", repr(exp)[:300], "\n   seq      ", repr(seq)[:300], "\n   tp       ", repr(tp)[:300], "\n   oracle   ", repr(orc)[:300])
    print(f"seed={seed} cases={n} compared={done} skipped={skipped} handed_over={handed} error_messages_compared={worded} bad={bad} time={time.time() - t0:.1f}s")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
