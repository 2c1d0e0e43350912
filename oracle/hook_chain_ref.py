"""Restatement of the three pattern plugins' data paths with CPython's `re` — the same matcher the
reference uses (stdlib, present on the GPU box; /root/reference is not).

  regex_filter   plugins/regex_filter/search_replace.py:59-75 (compile), :115-132 (tool_pre_invoke),
                 :134-156 (tool_post_invoke)
  deny_filter    plugins/deny_filter/deny.py:36-46 (config), :48-69 (prompt_pre_fetch)
  harmful        plugins/harmful_content_detector/harmful_content_detector.py:36-52 (lexicons),
                 :70-87 (compile), :92-107 (_scan_text), :110-139 (_iter_strings), :183-213 (hook)
"""
from __future__ import annotations

import re
from typing import Any, Dict, Iterable, List, Optional, Tuple

# harmful_content_detector.py:36-52
DEFAULT_LEXICONS: Dict[str, List[str]] = {
    "self_harm": [r"\bkill myself\b", r"\bsuicide\b", r"\bself-harm\b", r"\bwant to die\b"],
    "violence": [r"\bkill (?:him|her|them|someone)\b", r"\bshoot (?:him|her|them|someone)\b", r"\bstab (?:him|her|them|someone)\b"],
    "hate": [r"\b(?:kill|eradicate) (?:[a-z]+) people\b", r"\b(?:racial slur|hate speech)\b"],
}
DEFAULT_BLOCK_ON = ["self_harm", "violence", "hate"]  # harmful_content_detector.py:65


# ---------------------------------------------------------------- regex_filter
def regex_compile_rules(words: List[Dict[str, str]]) -> List[Tuple["re.Pattern[str]", str]]:
    """search_replace.py:67-75 — invalid patterns are silently skipped."""
    out = []
    for w in words:
        try:
            out.append((re.compile(w["search"]), w["replace"]))
        except re.error:
            pass
    return out


def regex_apply_dict(rules, d: Dict[str, Any]) -> Dict[str, Any]:
    """search_replace.py:125-130 / :145-150 — rule-major order over top-level str values."""
    out = dict(d)
    for pattern, replacement in rules:
        for k, v in out.items():
            if isinstance(v, str):
                out[k] = pattern.sub(replacement, v)
    return out


def regex_apply_str(rules, s: str) -> str:
    """search_replace.py:151-155."""
    for pattern, replacement in rules:
        s = pattern.sub(replacement, s)
    return s


# ---------------------------------------------------------------- deny_filter
def deny_first_hit(words: List[str], args: Optional[Dict[str, Any]]) -> Optional[str]:
    """deny.py:58-68 — first key (dict order) whose value contains any deny word, else None."""
    if args:
        for key in args:
            if any(word in args[key] for word in words):
                return key
    return None


# ---------------------------------------------------------------- harmful_content_detector
def harmful_compile(categories: Optional[Dict[str, List[str]]] = None) -> Dict[str, List["re.Pattern[str]"]]:
    """harmful_content_detector.py:70-87."""
    src = categories if categories is not None else DEFAULT_LEXICONS
    return {cat: [re.compile(p, re.IGNORECASE) if isinstance(p, str) else p for p in pats] for cat, pats in src.items()}


def harmful_scan_text(text: str, cats) -> List[Tuple[str, str]]:
    """harmful_content_detector.py:92-107."""
    findings = []
    for cat, pats in cats.items():
        for pat in pats:
            if pat.search(text):
                findings.append((cat, pat.pattern))
    return findings


def iter_strings(value: Any) -> Iterable[Tuple[str, str]]:
    """harmful_content_detector.py:110-139."""
    def walk(obj, path):
        if isinstance(obj, str):
            yield path, obj
        elif isinstance(obj, dict):
            for k, v in obj.items():
                yield from walk(v, f"{path}.{k}" if path else str(k))
        elif isinstance(obj, list):
            for i, v in enumerate(obj):
                yield from walk(v, f"{path}[{i}]")
    yield from walk(value, "")


def harmful_tool_post(result: Any, cats, block_on=DEFAULT_BLOCK_ON) -> Dict[str, Any]:
    """harmful_content_detector.py:183-213 — returns the fields of the hook's PluginResult."""
    if isinstance(result, (dict, list)):
        findings: List[Tuple[str, str]] = []
        for _, s in iter_strings(result):
            findings.extend(harmful_scan_text(s, cats))
    elif isinstance(result, str):
        findings = harmful_scan_text(result, cats)
    else:
        findings = []
    c = sorted(set(x for x, _ in findings))
    if any(x in block_on for x in c):
        return {"continue_processing": False, "violation": {"reason": "Harmful content", "description": f"Detected categories: {', '.join(c)}",
                                                            "code": "HARMFUL_CONTENT", "details": {"categories": c, "findings": findings[:5]}}, "metadata": {}}
    return {"continue_processing": True, "violation": None, "metadata": {"harmful_categories": c} if c else {}}


# ---------------------------------------------------------------- the fused scan, per unit
def scan_bitmaps(units: List[str], search_patterns: List[Tuple[str, int]], literals: List[str], sub_patterns: List[Tuple[str, int]]) -> List[int]:
    """Bitmap per unit in the order the product assigns bits: search patterns, literals, sub rules.
    Equivalent to running `pat.search(unit)` / `word in unit` for each (the three plugins' inner
    tests, cited above) — used as the checker for the fused GPU scan."""
    comp = [re.compile(p, f) for p, f in search_patterns]
    subs = [re.compile(p, f) for p, f in sub_patterns]
    out = []
    for u in units:
        v = 0
        bit = 0
        for c in comp:
            if c.search(u):
                v |= 1 << bit
            bit += 1
        for w in literals:
            if w in u:
                v |= 1 << bit
            bit += 1
        for c in subs:
            if c.search(u):
                v |= 1 << bit
            bit += 1
        out.append(v)
    return out
