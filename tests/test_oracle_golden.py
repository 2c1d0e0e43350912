"""Pins the oracle (oracle/hook_chain_ref.py) against golden vectors produced by the reference's own
plugin files (tools/gen_golden.py ran them in the build container; /root/reference is absent on the
GPU box).  CPU only."""
import json
import os

import pytest

from oracle import hook_chain_ref as ref

GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def gold():
    with open(os.path.join(GOLD, "pattern_plugins.json"), encoding="utf-8") as f:
        return json.load(f)


def test_regex_filter_golden(gold):
    n = 0
    for block in gold["regex_filter"]:
        rules = ref.regex_compile_rules(block["words"])
        for c in block["cases"]:
            if c["hook"] == "tool_pre_invoke":
                got = ref.regex_apply_dict(rules, c["args"]) if c["args"] else c["args"]
                assert got == c["out_args"]
            else:
                res = c["result"]
                if res and isinstance(res, dict):
                    got = ref.regex_apply_dict(rules, res)
                elif res and isinstance(res, str):
                    got = ref.regex_apply_str(rules, res)
                else:
                    got = res
                assert got == c["out_result"]
            n += 1
    assert n >= 200


def test_regex_filter_templates_golden():
    """Rules that can match "" and templates with group references (the reference hands `replace` to `pattern.sub` as is)."""
    with open(os.path.join(GOLD, "regex_filter_templates.json"), encoding="utf-8") as f:
        blocks = json.load(f)
    n = 0
    for block in blocks:
        rules = ref.regex_compile_rules(block["words"])
        for c in block["cases"]:
            if c["hook"] == "tool_pre_invoke":
                assert (ref.regex_apply_dict(rules, c["args"]) if c["args"] else c["args"]) == c["out_args"]
            else:
                res = c["result"]
                got = ref.regex_apply_dict(rules, res) if res and isinstance(res, dict) else ref.regex_apply_str(rules, res) if res and isinstance(res, str) else res
                assert got == c["out_result"]
            n += 1
    assert n >= 300


def test_deny_filter_golden(gold):
    n = 0
    for block in gold["deny_filter"]:
        for c in block["cases"]:
            hit = ref.deny_first_hit(block["words"], c["args"])
            assert (hit is not None) == c["blocked"]
            if c["blocked"]:
                assert c["violation"]["code"] == "deny" and c["violation"]["reason"] == "Prompt not allowed"
            n += 1
    assert n >= 150


def test_harmful_golden(gold):
    n = blocked = 0
    for block in gold["harmful"]:
        cfg = block["config"] or {}
        cats = ref.harmful_compile(cfg.get("categories"))
        block_on = cfg.get("block_on", ref.DEFAULT_BLOCK_ON)
        for c in block["cases"]:
            got = ref.harmful_tool_post(c["result"], cats, block_on)
            assert got["continue_processing"] == c["continue_processing"]
            assert got["metadata"] == (c["metadata"] or {})
            if c["violation"]:
                v = got["violation"]
                assert v["description"] == c["violation"]["description"]
                assert v["details"]["categories"] == c["violation"]["details"]["categories"]
                assert [list(x) for x in v["details"]["findings"]] == [list(x) for x in c["violation"]["details"]["findings"]]
                blocked += 1
            else:
                assert got["violation"] is None
            n += 1
    assert n >= 200 and blocked >= 20
