# -*- coding: utf-8 -*-
"""Regex front-end: Python `re` pattern text -> serialized AST for the C++/CUDA back-end.

The reference's pattern plugins hand their patterns to CPython's `re`:
  * regex_filter        re.compile(word.search)            /root/reference/plugins/regex_filter/search_replace.py:71
  * harmful_content     re.compile(p, re.IGNORECASE)       /root/reference/plugins/harmful_content_detector/harmful_content_detector.py:76,85
  * deny_filter         `word in value` (== escaped literal) /root/reference/plugins/deny_filter/deny.py:59-60

To be bit-exact with that matcher we do not re-implement its syntax or its Unicode tables.
The pattern is parsed by CPython's own sre parser (`re._parser.parse`), and every single-character
node (LITERAL / NOT_LITERAL / IN / ANY, with whatever IGNORECASE / ASCII / DOTALL flags are in
scope) is resolved to the explicit set of code points it matches by compiling *that node alone*
with sre and scanning all 0x110000 code points once (cached).  The back-end therefore receives
only interval sets, ordered alternations, repeats and zero-width assertions, and the result is
exact for whichever CPython version hosts the gateway.

Constructs a DFA cannot express (back-references, look-around, atomic groups, possessive
repeats, conditional groups, LOCALE) raise `UnsupportedPattern`; the plugins surface that loudly
at initialisation — there is no CPU fallback.
"""
from __future__ import annotations

import re
import re._compiler as _compiler  # type: ignore[import]
import re._constants as _k  # type: ignore[import]
import re._parser as _parser  # type: ignore[import]
from functools import lru_cache
from typing import List, Sequence, Tuple, Union

# AST opcodes — keep in sync with csrc/re_backend.h
A_EMPTY, A_SET, A_CAT, A_ALT, A_REPEAT, A_ASSERT, A_GROUP = 0, 1, 2, 3, 4, 5, 6
AS_WORD_B, AS_NOT_WORD_B, AS_BEGIN_STRING, AS_BEGIN_LINE, AS_END_STRING, AS_END_LINE, AS_END_DOLLAR = 1, 2, 3, 4, 5, 6, 7
REPEAT_INF = 0xFFFFFFFF
MAX_CP = 0x10FFFF

Ranges = Tuple[Tuple[int, int], ...]


class UnsupportedPattern(ValueError):
    """The pattern is valid Python `re` syntax but cannot run on the GPU engine."""


@lru_cache(maxsize=1)
def _all_chars() -> str:
    return "".join(map(chr, range(MAX_CP + 1)))


def _scan_ranges(pat: "re.Pattern[str]") -> Ranges:
    return tuple((m.start(), m.end() - 1) for m in pat.finditer(_all_chars()))


@lru_cache(maxsize=None)
def word_set() -> Ranges:
    r"""Code points that are `\w` for str patterns (used for `\b` / `\B` contexts)."""
    return _scan_ranges(re.compile(r"\w+"))


_FLAG_MASK = _k.SRE_FLAG_IGNORECASE | _k.SRE_FLAG_ASCII | _k.SRE_FLAG_UNICODE | _k.SRE_FLAG_DOTALL


_atom_cache: dict = {}


def _atom_ranges(node, flags: int) -> Ranges:
    """Exact set of code points matched by one single-character sre node under `flags`."""
    flags &= _FLAG_MASK
    op, av = node
    # cheap exact shortcuts (no engine scan needed)
    if op is _k.LITERAL and not (flags & _k.SRE_FLAG_IGNORECASE):
        return ((av, av),)
    if op is _k.ANY:
        if flags & _k.SRE_FLAG_DOTALL:
            return ((0, MAX_CP),)
        return ((0, 9), (11, MAX_CP))
    key = (repr(node), flags)
    hit = _atom_cache.get(key)
    if hit is not None:
        return hit
    state = _parser.State()
    state.flags = flags
    state.str = ""
    inner = _parser.SubPattern(state, [node])
    outer = _parser.SubPattern(state, [(_k.MAX_REPEAT, (1, _k.MAXREPEAT, inner))])
    compiled = _compiler.compile(outer, flags)
    out = _scan_ranges(compiled)
    _atom_cache[key] = out
    return out


def _emit_set(out: List[int], ranges: Ranges) -> None:
    out.append(A_SET)
    out.append(len(ranges))
    for lo, hi in ranges:
        out.append(lo)
        out.append(hi)


def _combine_flags(flags: int, add: int, delete: int) -> int:
    # mirrors re._compiler._combine_flags
    if add & _parser.TYPE_FLAGS:
        flags &= ~_parser.TYPE_FLAGS
    return (flags | add) & ~delete


def _emit_seq(out: List[int], items: Sequence, flags: int, mode: str, tail: bool, groups: bool = False) -> None:
    items = list(items)
    out.append(A_CAT)
    out.append(len(items))
    for i, node in enumerate(items):
        _emit_node(out, node, flags, mode, tail and i == len(items) - 1, groups)


def _emit_node(out: List[int], node, flags: int, mode: str, tail: bool, groups: bool = False) -> None:
    op, av = node
    if op in (_k.LITERAL, _k.NOT_LITERAL, _k.IN, _k.ANY):
        ranges = _atom_ranges(node, flags)
        _emit_set(out, ranges)  # an empty set is legal: the node then never matches
    elif op is _k.BRANCH:
        _, alts = av
        out.append(A_ALT)
        out.append(len(alts))
        for alt in alts:
            _emit_seq(out, alt, flags, mode, tail, groups)
    elif op is _k.SUBPATTERN:
        _group, add, delete, sub = av
        if groups and _group is not None:
            out.extend((A_GROUP, _group))          # capture spans for replacement templates with group references
        _emit_seq(out, sub, _combine_flags(flags, add, delete), mode, tail, groups)
    elif op in (_k.MAX_REPEAT, _k.MIN_REPEAT):
        # sre ends an unbounded loop on a zero-width iteration; the priority closure of the ordered automaton would go on into
        # the body's lower-priority branch instead (r'(?:c)+(?:(?:s)*?)+' and friends).  Existence (search) is unaffected.
        # Bounded loops share the rule (an empty iteration ends the loop once `min` is reached), which decides extents and group spans.
        if mode == "sub" and av[1] > 1 and av[2].getwidth()[0] == 0:
            raise UnsupportedPattern("repeat of a sub-pattern that can match the empty string (substitution mode)")
        mn, mx, sub = av
        out.extend((A_REPEAT, mn, REPEAT_INF if mx == _k.MAXREPEAT else mx, 1 if op is _k.MAX_REPEAT else 0))
        _emit_seq(out, sub, flags, mode, False, groups)
    elif op is _k.AT:
        if flags & _k.SRE_FLAG_LOCALE:
            raise UnsupportedPattern("LOCALE flag")
        multiline = bool(flags & _k.SRE_FLAG_MULTILINE)
        if av is _k.AT_BEGINNING:
            out.extend((A_ASSERT, AS_BEGIN_LINE if multiline else AS_BEGIN_STRING))
        elif av is _k.AT_BEGINNING_STRING:
            out.extend((A_ASSERT, AS_BEGIN_STRING))
        elif av is _k.AT_END_STRING:
            out.extend((A_ASSERT, AS_END_STRING))
        elif av is _k.AT_END:
            if multiline:
                out.extend((A_ASSERT, AS_END_LINE))
            else:
                # `$` == end of string, or just before a final "\n".  At the end of a SEARCH pattern that is (?:\Z|\n\Z) — exact for
                # existence, and it keeps the automaton free of the extra character class.  Anywhere else, and in substitution rules
                # (where consuming the newline would change the extent), it is the assertion AS_END_DOLLAR: the back-end gives "the
                # newline that is the unit's last character" a class of its own and the assertion looks one character ahead.
                if mode == "search" and tail:
                    out.extend((A_ALT, 2, A_ASSERT, AS_END_STRING, A_CAT, 2, A_SET, 1, 10, 10, A_ASSERT, AS_END_STRING))
                else:
                    out.extend((A_ASSERT, AS_END_DOLLAR))
        elif av in (_k.AT_BOUNDARY, _k.AT_NON_BOUNDARY):
            if flags & _k.SRE_FLAG_ASCII:
                raise UnsupportedPattern(r"ASCII-mode \b")
            out.extend((A_ASSERT, AS_WORD_B if av is _k.AT_BOUNDARY else AS_NOT_WORD_B))
        else:
            raise UnsupportedPattern(f"assertion {av}")
    elif op in (_k.ASSERT, _k.ASSERT_NOT):
        raise UnsupportedPattern("look-ahead / look-behind")
    elif op in (_k.GROUPREF, _k.GROUPREF_EXISTS):
        raise UnsupportedPattern("back-reference")
    elif op is _k.ATOMIC_GROUP:
        raise UnsupportedPattern("atomic group")
    elif op is _k.POSSESSIVE_REPEAT:
        raise UnsupportedPattern("possessive repeat")
    elif op is _k.FAILURE:
        raise UnsupportedPattern("(?!) failure node")
    else:
        raise UnsupportedPattern(f"sre node {op}")


def compile_ast(pattern: str, flags: int = 0, mode: str = "search", groups: bool = False) -> List[int]:
    """Parse `pattern` exactly as `re.compile(pattern, flags)` would and serialize it.

    mode: "search" (existence, any match)   or   "sub" (leftmost-first extents).
    Raises `re.error` for invalid patterns (same as the reference would see) and
    `UnsupportedPattern` for constructs the GPU engine cannot express.
    """
    if not isinstance(pattern, str):
        raise UnsupportedPattern("bytes patterns")
    parsed = _parser.parse(pattern, flags)
    final_flags = parsed.state.flags | flags
    if final_flags & _k.SRE_FLAG_LOCALE:
        raise UnsupportedPattern("LOCALE flag")
    if not (final_flags & _k.SRE_FLAG_ASCII):
        final_flags |= _k.SRE_FLAG_UNICODE
    out: List[int] = []
    _emit_seq(out, parsed, final_flags, mode, True, groups)
    return out


def literal_ast(text: str) -> List[int]:
    """AST of a case-sensitive literal substring (deny_filter's `word in value`)."""
    out: List[int] = [A_CAT, len(text)]
    for ch in text:
        out.extend((A_SET, 1, ord(ch), ord(ch)))
    return out


def template_parts(template: str, pattern: "re.Pattern[str]") -> List[Union[str, int]]:
    """A `re.sub` replacement template as a flat list of literal strings and group indices, parsed by sre's own template parser
    (escape processing, `\\1`, `\\g<name>`, `\\g<0>`; bad templates raise `re.error` exactly like `pattern.sub` would)."""
    parts = _parser.parse_template(template, pattern)
    if isinstance(parts, tuple):
        # Python 3.11 (the reference supports >= 3.11): (groups, literals) — literals[i] is None where groups say (i, group)
        groups, literals = parts
        lits = list(literals)
        for i, g in groups:
            lits[i] = g
        parts = lits
    return [p for p in parts if p is not None and p != ""]


def encode_template(parts: Sequence[Union[str, int]]):
    """-> (literal bytes, uint32 triples {kind, a, b}) for cf_builder_set_template."""
    lit = bytearray()
    out: List[int] = []
    for p in parts:
        if isinstance(p, int):
            out.extend((1, p, 0))
        else:
            b = p.encode("utf-8", "surrogatepass")
            out.extend((0, len(lit), len(b)))
            lit += b
    return bytes(lit), out
