"""GPU: the json_repair drop-in (parse status from the JSON kernel, trailing-comma rule through sub_kernel) against the golden vectors recorded
from the reference's plugin, stand-alone and inside BatchedPluginManager (== the sequential executor).  (File name: runs late in the suite.)"""
import json

import pytest

import test_json_repair as tj

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def cases():
    with open(tj.GOLD, encoding="utf-8") as f:
        return json.load(f)["cases"]


def test_dropin_matches_reference_golden_gpu(cases):
    from mcp_context_forge_b200.plugins.json_repair import JSONRepairPlugin

    plug = tj.check_dropin(cases, JSONRepairPlugin)
    assert plug.unsupported_kept == 0


@pytest.mark.parametrize("regex_prio", [50, 150])
def test_dropin_inside_the_batched_manager_gpu(cases, regex_prio):
    tj.check_in_manager(cases, regex_prio)
