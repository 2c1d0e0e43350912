"""The masking MODULE (request_logging_masking_native_extension -> mcp_context_forge_b200/masking.py: key collection, classification of a tree's
key names in one launch, cookie splitting, header walk, bytes API, the middleware's fallback probes) with its launches on the CPU simulator
(tests/hostsim_batcher.py: host build of csrc/json_mask.h and of the scan tables): the very test bodies of tests/test_mask_gpu.py — the crate's
unit tests and benchmark vectors, the twin's golden vectors, the oracle — run in the `not gpu` suite as well."""
import importlib
import json

import pytest

import hostsim_batcher
import test_mask_gpu as tg


@pytest.fixture()
def mod(monkeypatch):
    hostsim_batcher.install(monkeypatch)
    return importlib.import_module("request_logging_masking_native_extension")


@pytest.fixture(scope="module")
def gold():
    with open(tg.GOLD, encoding="utf-8") as f:
        return json.load(f)


def test_crate_unit_tests_and_benchmark_parity_vectors(mod):
    tg.test_crate_unit_tests_and_benchmark_parity_vectors(mod)


def test_twin_golden_object_api(mod, gold):
    tg.test_twin_golden_object_api(mod, gold)


def test_benchmark_scenario_payload(mod):
    tg.test_benchmark_scenario_payload(mod)


def test_non_json_fallback_and_header_batch(mod):
    tg.test_non_json_fallback_and_header_batch()
