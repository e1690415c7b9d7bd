import importlib
import os
import sys
import threading

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


def load_pkg():
    """The package directory is `nav-slam_b200` (hyphen): import it by name."""
    return importlib.import_module("nav-slam_b200")


def big_stack(fn, *a, **kw):
    """Run fn on a thread with a 512 MB stack: the reference keeps whole clouds on the
    stack (6.8 MB of locals in slam_localization at 64x2048)."""
    box = {}

    def run():
        try:
            box["v"] = fn(*a, **kw)
        except BaseException as e:  # noqa: BLE001
            box["e"] = e

    old = threading.stack_size(512 * 1024 * 1024)
    try:
        t = threading.Thread(target=run)
        t.start()
        t.join()
    finally:
        threading.stack_size(old)
    if "e" in box:
        raise box["e"]
    return box.get("v")


@pytest.fixture(scope="session")
def pkg():
    return load_pkg()


@pytest.fixture(scope="session")
def synth():
    return importlib.import_module("nav-slam_b200.synth")


@pytest.fixture(scope="session")
def oracle():
    from oracle_lib import Oracle
    return Oracle()
