"""TEST INFRASTRUCTURE ONLY (the checker, never the product path): restatement of
/root/reference/plugins/sql_sanitizer/sql_sanitizer.py on CPython's `re`.

  patterns             sql_sanitizer.py:33-44
  strip_comments       :102-114   (line comments first, then block comments)
  has_interpolation    :117-128   (note the precedence: `a or b or (c and d)`)
  find_issues          :131-163
  scan_args            :166-222   (issue labels `key` / `key[]`, `scanned` keyed by the innermost key)
  hook                 :239-289   (block -> violation; else stripped args merged over the old ones; else issue metadata)

Pinned by tests/golden/sql_sanitizer.json, recorded from the reference's own file (tools/gen_golden.py).
"""
from __future__ import annotations

import re
from typing import Any, Dict, List, Optional, Tuple

DEFAULT_BLOCKED = [r"\bDROP\b", r"\bTRUNCATE\b", r"\bALTER\b", r"\bGRANT\b", r"\bREVOKE\b"]
LINE_COMMENT = re.compile(r"--.*?$", re.MULTILINE)
BLOCK_COMMENT = re.compile(r"/\*.*?\*/", re.DOTALL)
DELETE_FROM = re.compile(r"\bDELETE\b\s+\bFROM\b", re.IGNORECASE)
UPDATE = re.compile(r"\bUPDATE\b\s+\w+", re.IGNORECASE)
WHERE = re.compile(r"\bWHERE\b", re.IGNORECASE)
DEFAULTS = {"fields": None, "blocked_statements": DEFAULT_BLOCKED, "block_delete_without_where": True, "block_update_without_where": True,
            "strip_comments": True, "require_parameterization": False, "block_on_violation": True}


def config(cfg: Optional[Dict[str, Any]]) -> Dict[str, Any]:
    c = dict(DEFAULTS)
    c.update(cfg or {})
    c["blocked_statements"] = [re.compile(p, re.IGNORECASE) if isinstance(p, str) else p for p in c["blocked_statements"]]
    return c


def strip_comments(sql: str) -> str:
    return BLOCK_COMMENT.sub("", LINE_COMMENT.sub("", sql))


def has_interpolation(sql: str) -> bool:
    return "+" in sql or "%." in sql or ("{" in sql and "}" in sql)


def find_issues(sql: str, c: Dict[str, Any]) -> List[str]:
    original = sql
    if c["strip_comments"]:
        sql = strip_comments(sql)
    issues = [f"Blocked statement matched: {p.pattern}" for p in c["blocked_statements"] if p.search(sql)]
    if c["block_delete_without_where"] and DELETE_FROM.search(sql) and not WHERE.search(sql):
        issues.append("DELETE without WHERE clause")
    if c["block_update_without_where"] and UPDATE.search(sql) and not WHERE.search(sql):
        issues.append("UPDATE without WHERE clause")
    if c["require_parameterization"] and has_interpolation(original):
        issues.append("Possible non-parameterized interpolation detected")
    return issues


def scan_args(args: Optional[Dict[str, Any]], c: Dict[str, Any]) -> Tuple[List[str], Dict[str, Any]]:
    issues: List[str] = []
    scanned: Dict[str, Any] = {}
    fields = c["fields"]

    def visit(key, value):
        if isinstance(value, str):
            if fields is None or key in fields:
                issues.extend(f"{key}: {m}" for m in find_issues(value, c))
                if c["strip_comments"]:
                    clean = strip_comments(value)
                    if clean != value:
                        scanned[key] = clean
        elif isinstance(value, dict):
            for k, v in value.items():
                visit(k, v)
        elif isinstance(value, list):
            for item in value:
                if isinstance(item, dict):
                    for k, v in item.items():
                        visit(k, v)
                elif isinstance(item, str) and (fields is None or key in fields):
                    issues.extend(f"{key}[]: {m}" for m in find_issues(item, c))

    for k, v in (args or {}).items():
        visit(k, v)
    return issues, scanned


def hook(args: Optional[Dict[str, Any]], c: Dict[str, Any], where: str) -> Dict[str, Any]:
    """-> {continue_processing, violation (dict | None), out_args (dict | None), metadata}; where = 'tool args' | 'prompt args'."""
    issues, scanned = scan_args(args or {}, c)
    if issues and c["block_on_violation"]:
        return {"continue_processing": False, "out_args": None, "metadata": {},
                "violation": {"reason": "Risky SQL detected", "description": f"Potentially dangerous SQL detected in {where}", "code": "SQL_SANITIZER", "details": {"issues": issues}}}
    if scanned:
        return {"continue_processing": True, "violation": None, "out_args": {**(args or {}), **scanned}, "metadata": {"sql_sanitized": True}}
    return {"continue_processing": True, "violation": None, "out_args": None, "metadata": {"sql_issues": issues} if issues else {}}
