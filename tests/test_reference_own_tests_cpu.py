"""Container-only: the reference's OWN test files, unmodified, against this repo — the executor restatement (`cpex.framework` surface, agent
plugin chains) and the drop-in plugins on the engine's CPU simulator.  Skipped where /root/reference does not exist (the GPU box): the
travelling form of the same evidence is tests/golden/reference_tests_run_*.json + tests/test_executor_reference_cases.py."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.skipif(not os.path.isdir("/root/reference/tests"), reason="/root/reference is not on this machine")


def _run(tmp_path, *args):
    out = tmp_path / "run.json"
    p = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "run_reference_tests.py"), "--tb=line", "-q", "--json", str(out), *args],
                       capture_output=True, text=True, timeout=900)
    return p, json.loads(out.read_text())


def test_reference_executor_tests_pass_on_the_restated_executor(tmp_path):
    """tests/acceptance/plugins/test_cpex_contract.py + tests/unit/mcpgateway/plugins/agent/test_agent_plugins.py (default targets)."""
    p, s = _run(tmp_path)
    other = {k: v for k, v in s["not_passed_detail"].items() if v["bucket"] == "other"}
    assert not other, p.stdout[-3000:]
    assert s["passed"] >= 27 and s["per_file"]["tests/unit/mcpgateway/plugins/agent/test_agent_plugins.py"] == {"passed": 8}


def test_reference_plugin_tests_pass_on_the_drop_ins(tmp_path):
    """tests/unit/plugins/toon_encoder/test_toon_encoder.py, tests/unit/plugins/test_sql_sanitizer.py,
    tests/unit/mcpgateway/plugins/plugins/code_safety_linter/test_code_safety_linter.py with `plugins.<x>` -> this repo's drop-ins."""
    p, s = _run(tmp_path, "--dropin")
    assert p.returncode == 0 and s["not_passed"] == 0 and s["passed"] >= 23, p.stdout[-3000:]


def test_committed_run_records_have_no_unexplained_failure():
    for name in ("framework", "dropin"):
        with open(os.path.join(ROOT, "tests", "golden", f"reference_tests_run_{name}.json")) as f:
            s = json.load(f)
        assert s["passed"] > 0 and all(v["bucket"] != "other" for v in s["not_passed_detail"].values())


@pytest.mark.parametrize("tool,args", [("fuzz_vs_reference.py", ["5", "4000"]), ("fuzz_mask_vs_reference.py", ["5", "8000", "800"]), ("fuzz_json_repair_vs_reference.py", ["5", "4000"]),
                                       ("fuzz_plugins_vs_reference.py", ["5", "12"]), ("fuzz_chain_vs_reference.py", ["5", "8", "40"])])
def test_live_differential_fuzz_against_the_reference(tool, args):
    """A short run of each container-only differential fuzzer (kernel source on the host build / warp emulator, drop-ins and the batched chain on
    the CPU simulator) against the reference's own modules; the long runs of the round are recorded in DESIGN.md §5."""
    p = subprocess.run([sys.executable, os.path.join(ROOT, "tools", tool), *args], capture_output=True, text=True, timeout=900)
    assert p.returncode == 0 and " bad=0" in p.stdout, (p.stdout[-2000:], p.stderr[-2000:])
