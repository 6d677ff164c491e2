import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def golden(name: str, line: int):
    """Literal list that starts at `line` of /root/reference/test/<name>.jl."""
    import numpy as np

    with open(os.path.join(GOLDEN, name + ".json")) as f:
        data = json.load(f)
    for item in data["lists"]:
        if item["line"] == line:
            return np.array(item["values"], dtype=np.float64)
    raise KeyError(f"{name}.jl has no numeric list starting at line {line}")


@pytest.fixture(scope="session")
def orc():
    from oracle import oracle

    oracle.build()
    return oracle


@pytest.fixture(scope="session")
def gp():
    """The product package (gempic.jl_b200/), loaded through __graft_entry__.load_package()."""
    import __graft_entry__ as ge

    return ge.load_package()
