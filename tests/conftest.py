import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def golden():
    with open(os.path.join(ROOT, "tests", "golden", "reference_vectors.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def ctx():
    import acvm_b200
    c = acvm_b200.Context(0)
    yield c
    c.close()


@pytest.fixture(scope="session")
def pctx():
    """Context that opted in to the Pedersen kernel (parity with barretenberg unpinned: refused by default)."""
    import acvm_b200
    c = acvm_b200.Context(0)
    c.set_option("pedersen_unpinned", 1)
    yield c
    c.close()


def inputs_to_dicts(inp: bytes, batch: int, input_witnesses):
    n = len(input_witnesses)
    return [{w: int.from_bytes(inp[(i * n + k) * 32:(i * n + k + 1) * 32], "big") for k, w in enumerate(input_witnesses)}
            for i in range(batch)]


def witness_rows(out: bytes, batch: int, n_out: int):
    return [[int.from_bytes(out[(i * n_out + k) * 32:(i * n_out + k + 1) * 32], "big") for k in range(n_out)] for i in range(batch)]
