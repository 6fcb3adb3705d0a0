import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def goldens():
    with open(os.path.join(ROOT, "tests", "golden", "reference_goldens.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def native_built():
    """Builds the native libraries once (nvcc cross-compiles without a GPU)."""
    import kripke_b200
    from oracle import ko
    paths = kripke_b200.lib_paths()
    if not all(os.path.exists(p) for p in paths):
        kripke_b200.build()
    ko.build()
    return paths


@pytest.fixture(scope="session")
def gpu(native_built):
    """Binds the process to cuda:0; GPU tests must run the CUDA path or fail -- never skip silently."""
    import kripke_b200
    kripke_b200.init_device(0)
    return kripke_b200


def parse_args(args):
    """kripke command line -> kwargs of oracle.ko.Problem"""
    a = args.split()
    kw = {}
    i = 0
    while i < len(a):
        o, v = a[i], a[i + 1] if i + 1 < len(a) else None
        if o == "--zones": kw["zones"] = tuple(map(int, v.split(",")))
        elif o == "--zset": kw["zset"] = tuple(map(int, v.split(",")))
        elif o == "--procs": kw["procs"] = tuple(map(int, v.split(",")))
        elif o == "--groups": kw["groups"] = int(v)
        elif o == "--quad": kw["quad"] = tuple(map(int, v.split(":"))) if ":" in v else int(v)
        elif o == "--legendre": kw["legendre"] = int(v)
        elif o == "--layout": kw["layout"] = v
        elif o == "--gset": kw["gset"] = int(v)
        elif o == "--dset": kw["dset"] = int(v)
        elif o == "--sigt": kw["sigt"] = tuple(map(float, v.split(",")))
        elif o == "--sigs": kw["sigs"] = tuple(map(float, v.split(",")))
        elif o == "--pmethod": kw["_bj"] = (v.lower() == "bj")
        elif o == "--niter": kw["_niter"] = int(v)
        i += 2
    return kw


def oracle_problem(args):
    from oracle import ko
    kw = parse_args(args)
    niter = kw.pop("_niter", 10)
    bj = kw.pop("_bj", False)
    return ko.Problem(**kw), niter, bj
