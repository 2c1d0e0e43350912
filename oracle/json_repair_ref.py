"""TEST INFRASTRUCTURE ONLY: restatement of /root/reference/plugins/json_repair/json_repair.py on CPython `json` / `re`
(`_try_parse` :36-50 = "orjson.loads succeeds", `_repair` :53-78 — single quotes when the text is bracketed and has no double quote,
trailing commas before a closing bracket, bare `key: value` wrapped in braces — and the hook :95-117).  orjson is stood in for by the
strict stdlib parser (no NaN / Infinity literals); the golden inputs keep clear of the remaining deltas (lone surrogate escapes, integers
beyond 64 bits, 1e400).  Pinned by tests/golden/json_repair.json, recorded from the reference's own file (tools/gen_golden.py json_repair)."""
from __future__ import annotations

import json
import re
from typing import Any, Dict, Optional

_BRACKETS = re.compile(r"^[\[{].*[\]}]$", flags=re.S)          # :32
_TRAILING_COMMA = re.compile(r",(\s*[}\]])")                    # :33


def parses(s: str) -> bool:
    try:
        json.loads(s, parse_constant=lambda c: (_ for _ in ()).throw(ValueError(c)))
        return True
    except (ValueError, RecursionError):
        return False


def candidates(s: str):
    """The texts `_repair` (:53-78) would try, in its order; a candidate that does not differ from its predecessor is not a candidate."""
    t = s.strip()
    base = t
    quoted = bool(_BRACKETS.match(t)) and "'" in t and '"' not in t
    if quoted:
        base = t.replace("'", '"')
        yield base
    without_commas = _TRAILING_COMMA.sub(r"\1", base)
    if without_commas != base:
        yield without_commas
    bare = not t.startswith("{") and ":" in t and "{" not in t and "}" not in t
    if bare:
        yield "{" + t + "}"


def repair(s: str) -> Optional[str]:
    return next((c for c in candidates(s) if parses(c)), None)


def hook(result: Any) -> Dict[str, Any]:
    """tool_post_invoke :95-117 as the golden file records it."""
    if isinstance(result, str) and not parses(result):
        fixed = repair(result)
        if fixed is not None:
            return {"continue_processing": True, "modified": True, "out_result": fixed, "metadata": {"repaired": True}}
    return {"continue_processing": True, "modified": False, "out_result": None, "metadata": {}}
