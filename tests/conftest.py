import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


@pytest.fixture(scope="session")
def golden():
    with open(os.path.join(ROOT, "tests", "golden", "golden.json")) as f:
        return json.load(f)["cases"]


@pytest.fixture(scope="session")
def golden_large():
    """Full-size cases (BASELINE configs 2-5) from the unmodified reference: tests/golden/make_golden_large.py."""
    with open(os.path.join(ROOT, "tests", "golden", "golden_large.json")) as f:
        return json.load(f)["cases"]


@pytest.fixture(scope="session")
def product_lib():
    """The built product library (compiles it if needed; nvcc cross-compiles without a GPU)."""
    from miniwfa_b200 import build
    build.build()
    from miniwfa_b200 import api
    return api.lib()


def case_inputs(case):
    from miniwfa_b200 import synth
    if "synth" in case:
        return synth.make_pair(*case["synth"])
    return case["t"].encode("latin-1"), case["q"].encode("latin-1")


@pytest.fixture(scope="session")
def golden_chain():
    with open(os.path.join(ROOT, "tests", "golden", "golden_chain.json")) as f:
        return json.load(f)["cases"]


def chain_case_inputs(case):
    """Inputs of a golden_chain.json case from its recipe (same builder as the generator script)."""
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    import make_golden_chain
    return make_golden_chain.build_inputs(case)
