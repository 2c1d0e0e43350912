"""TEST INFRASTRUCTURE ONLY: restatement of /root/reference/plugins/code_safety_linter/code_safety_linter.py on CPython `re`
(defaults :34-40, text selection :99-104, findings in pattern order :107-110, violation :111-121).
Pinned by tests/golden/code_safety.json, recorded from the reference's own file (tools/gen_golden.py)."""
from __future__ import annotations

import re
from typing import Any, Dict, List, Optional

DEFAULTS = [r"\beval\s*\(", r"\bexec\s*\(", r"\bos\.system\s*\(", r"\bsubprocess\.(Popen|call|run)\s*\(", r"\brm\s+-rf\b"]


def hook(result: Any, patterns: Optional[List[str]] = None) -> Dict[str, Any]:
    comp = [re.compile(p) for p in (DEFAULTS if patterns is None else patterns)]
    text = result if isinstance(result, str) else result.get("text") if isinstance(result, dict) and isinstance(result.get("text"), str) else None
    if not text:
        return {"continue_processing": True, "violation": None}
    findings = [c.pattern for c in comp if c.search(text)]
    if findings:
        return {"continue_processing": False, "violation": {"reason": "Unsafe code pattern", "description": "Detected unsafe code constructs", "code": "CODE_SAFETY",
                                                            "details": {"patterns": findings}}}
    return {"continue_processing": True, "violation": None}
