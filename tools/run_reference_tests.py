#!/usr/bin/env python
"""Run the REFERENCE'S OWN test files (unmodified, where they lie under /root/reference) against this repo's executor restatement
and drop-in plugins.  Container-only provenance tool: /root/reference does not exist on the GPU box, so nothing in `tests/` depends on
it; the outcome of a run is committed as `tests/golden/reference_tests_run.json` and the scenarios that pin the executor are restated
in `tests/test_executor_reference_cases.py` (each citing the reference test it follows).

What is swapped in (everything else is the reference's own code, imported from /root/reference):
  * `cpex.framework`                       -> mcp_context_forge_b200.cpex_compat  (cpex is third-party and absent from the tree)
  * `orjson`                               -> stand-in over stdlib json (as tools/gen_golden.py)
  * `mcpgateway.services` (package __init__) and `.logging_service` -> path-only package + stub logger (they pull in sqlalchemy)
  * with --dropin: `plugins.<x>.<y>`       -> the drop-in module of this repo for the §8 plugins (engine on the CPU simulator of the
                                              scan tables / the host build of the kernels' headers: tests/hostsim — no GPU here)
pytest-asyncio is not installed: a 15-line hook below runs `async def` tests on a fresh event loop.  The reference's conftest.py
files import the whole gateway (database, FastAPI app) and are skipped (`--noconftest`); the one thing they do for these files —
resetting the PluginManager's shared state between tests — is done by the hook.

usage: python tools/run_reference_tests.py [--dropin] [--json OUT] [pytest args / test paths relative to /root/reference]
"""
from __future__ import annotations

import argparse
import asyncio
import inspect
import json
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"

EXECUTOR_TESTS = [
    "tests/acceptance/plugins/test_cpex_contract.py",
    "tests/unit/mcpgateway/plugins/agent/test_agent_plugins.py",
]
DROPIN_TESTS = [
    "tests/unit/plugins/toon_encoder/test_toon_encoder.py",
    "tests/unit/plugins/test_sql_sanitizer.py",
    "tests/unit/mcpgateway/plugins/plugins/code_safety_linter/test_code_safety_linter.py",
    "tests/unit/mcpgateway/plugins/plugins/json_repair/test_json_repair.py",
]


def install_shims(dropin: bool) -> None:
    sys.path.insert(0, ROOT)
    from mcp_context_forge_b200.cpex_compat import install_as_cpex

    install_as_cpex(force=True)
    if "orjson" not in sys.modules:
        oj = types.ModuleType("orjson")

        class JSONDecodeError(ValueError):
            pass

        def loads(s):
            if isinstance(s, (bytes, bytearray, memoryview)):
                s = bytes(s).decode("utf-8")
            try:
                return json.loads(s, parse_constant=lambda c: (_ for _ in ()).throw(ValueError(c)))
            except ValueError as exc:
                raise JSONDecodeError(str(exc)) from exc

        oj.loads, oj.JSONDecodeError = loads, JSONDecodeError
        oj.dumps = lambda o, **_k: json.dumps(o, separators=(",", ":"), ensure_ascii=False).encode("utf-8")
        sys.modules["orjson"] = oj
    if REF not in sys.path:
        sys.path.insert(1, REF)
    # `mcpgateway.services/__init__.py` imports the whole gateway (sqlalchemy, the database): the package is registered by path only, and
    # the one module the reference's plugins need from it — logging_service (pythonjsonlogger, settings, log storage) — is a stub logger.
    import logging

    import mcpgateway  # noqa: F401  (the reference's own package)

    pkg = types.ModuleType("mcpgateway.services")
    pkg.__path__ = [os.path.join(REF, "mcpgateway", "services")]
    sys.modules["mcpgateway.services"] = pkg
    mcpgateway.services = pkg
    ls = types.ModuleType("mcpgateway.services.logging_service")

    class LoggingService:
        def get_logger(self, name):
            return logging.getLogger(name)

    ls.LoggingService = LoggingService
    sys.modules["mcpgateway.services.logging_service"] = ls
    pkg.logging_service = ls
    if dropin:
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import hostsim_batcher  # noqa: F401  (routes engine.Context / Program / run_batch to the CPU simulator; test infrastructure)

        import pytest

        hostsim_batcher.install(pytest.MonkeyPatch())
        import importlib

        import plugins  # the reference's package: only the leaf modules below are replaced

        for ref_mod, ours in (("plugins.toon_encoder.toon_encoder", "mcp_context_forge_b200.plugins.toon_encoder"),
                              ("plugins.sql_sanitizer.sql_sanitizer", "mcp_context_forge_b200.plugins.sql_sanitizer"),
                              ("plugins.code_safety_linter.code_safety_linter", "mcp_context_forge_b200.plugins.code_safety_linter"),
                              ("plugins.json_repair.json_repair", "mcp_context_forge_b200.plugins.json_repair"),
                              ("plugins.regex_filter.search_replace", "mcp_context_forge_b200.plugins.regex_filter"),
                              ("plugins.deny_filter.deny", "mcp_context_forge_b200.plugins.deny_filter"),
                              ("plugins.harmful_content_detector.harmful_content_detector", "mcp_context_forge_b200.plugins.harmful_content_detector")):
            importlib.import_module(ref_mod.rsplit(".", 1)[0])
            sys.modules[ref_mod] = importlib.import_module(ours)


class _Hook:
    def __init__(self):
        self.outcomes = {}
        self.reasons = {}

    @staticmethod
    def pytest_configure(config):
        config.addinivalue_line("markers", "asyncio: coroutine test (run by tools/run_reference_tests.py)")

    @staticmethod
    def pytest_runtest_setup(item):
        from cpex.framework import PluginManager

        reset = getattr(PluginManager, "reset", None)
        if reset is not None:
            reset()

    @staticmethod
    def pytest_pyfunc_call(pyfuncitem):
        fn = pyfuncitem.obj
        if inspect.iscoroutinefunction(fn):
            kw = {a: pyfuncitem.funcargs[a] for a in pyfuncitem._fixtureinfo.argnames}
            loop = asyncio.new_event_loop()
            try:
                loop.run_until_complete(fn(**kw))
            finally:
                loop.close()
            return True
        return None

    def pytest_runtest_logreport(self, report):
        if report.when == "call" or (report.when == "setup" and report.outcome != "passed"):
            self.outcomes[report.nodeid] = report.outcome
            if report.outcome == "failed":
                self.reasons[report.nodeid] = _last_line(report.longreprtext)

    def pytest_collectreport(self, report):
        if report.failed:
            self.outcomes[report.nodeid] = "collection error"
            self.reasons[report.nodeid] = _last_line(report.longreprtext)


def _last_line(text: str) -> str:
    lines = [ln.strip() for ln in text.splitlines() if ln.strip().startswith("E ")]
    return (lines[-1][2:].strip() if lines else text.strip().splitlines()[-1] if text.strip() else "")[:200]


def _why(reason: str) -> str:
    """Bucket of a failure: what is missing in THIS container (third-party packages the reference needs) vs. anything else."""
    for mod in ("sqlalchemy", "cpex_pii_filter", "cpex_rate_limiter", "cpex_secrets_detection", "cpex_encoded_exfil_detection", "cpex_retry_with_backoff",
                "cpex_url_reputation", "httpx_mock", "cpex.tools.models", "OPT_SORT_KEYS", "observability_adapter", "gateway_plugin_manager", "jwt", "hvac"):
        if mod in reason:
            return f"absent here: {mod}"
    return "other"


def main() -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument("--dropin", action="store_true")
    ap.add_argument("--json")
    ap.add_argument("rest", nargs="*")
    args, extra = ap.parse_known_args()
    if not os.path.isdir(REF):
        print("run_reference_tests: /root/reference is not here (container-only tool)")
        return 0
    install_shims(args.dropin)
    import pytest

    os.chdir(REF)                       # the reference's tests open their fixture YAMLs by relative path
    hook = _Hook()
    targets = args.rest or (DROPIN_TESTS if args.dropin else EXECUTOR_TESTS)
    rc = pytest.main(["-p", "no:cacheprovider", "--noconftest", "-c", os.devnull, "--rootdir", REF, "-q", *extra, *targets], plugins=[hook])
    if args.json:
        import collections

        per_file = collections.defaultdict(lambda: collections.Counter())
        for k, v in hook.outcomes.items():
            per_file[k.split("::")[0]][v] += 1
        not_passed = {k: {"outcome": v, "reason": hook.reasons.get(k, ""), "bucket": _why(hook.reasons.get(k, ""))}
                      for k, v in sorted(hook.outcomes.items()) if v not in ("passed", "skipped")}
        summary = {"what": "the reference's own test files, unmodified, run by tools/run_reference_tests.py against this repo (see its docstring for what is swapped in)",
                   "targets": targets, "dropin": args.dropin, "passed": sum(v == "passed" for v in hook.outcomes.values()),
                   "skipped": sum(v == "skipped" for v in hook.outcomes.values()), "not_passed": len(not_passed),
                   "not_passed_by_bucket": dict(collections.Counter(v["bucket"] for v in not_passed.values())),
                   "per_file": {k: dict(v) for k, v in sorted(per_file.items())}, "not_passed_detail": not_passed}
        with open(os.path.join(ROOT, args.json) if not os.path.isabs(args.json) else args.json, "w") as f:
            json.dump(summary, f, indent=1)
            f.write("\n")
    return int(rc)


if __name__ == "__main__":
    sys.exit(main())
