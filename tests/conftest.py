"""pytest configuration: `gpu` marker, shared fixtures (oracle bindings, golden vectors, the engine)."""
import importlib
import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu on the GPU box")
    config.addinivalue_line("markers", "slow: CPU test that takes more than ~20 s")


def load_cloud(name):
    return np.fromfile(os.path.join(GOLDEN, name), np.float32).reshape(-1, 3)


@pytest.fixture(scope="session")
def pkg():
    """The product package (directory name has hyphens, hence importlib)."""
    return importlib.import_module("cuda-go-icp_b200")


@pytest.fixture(scope="session")
def restated():
    from oracle.oracle import Restated
    return Restated()


@pytest.fixture(scope="session")
def reference():
    from oracle.oracle import Reference
    if not Reference.available():
        pytest.skip("oracle/_ref not built (needs /root/reference; `make -C oracle ref`)")
    return Reference()


@pytest.fixture(scope="session")
def small():
    return dict(np.load(os.path.join(GOLDEN, "small_vectors.npz")))


@pytest.fixture(scope="session")
def runs():
    return json.load(open(os.path.join(GOLDEN, "goicp_runs.json")))


@pytest.fixture(scope="session")
def bunny():
    return {"model": load_cloud("bunny_model_s0.1_seed1234.f32"), "data": load_cloud("bunny_data_s0.1_seed1235.f32"),
            "model_s": load_cloud("bunny_model_s0.033_seed1234.f32"), "data_s": load_cloud("bunny_data_s0.033_seed1235.f32")}


def rot_angle(Ra, Rb):
    """angle between two (nearly orthonormal float32) rotations, radians.  For small angles
    ||Ra - Rb||_F = 2*sqrt(2)*sin(theta/2); unlike arccos((tr-1)/2) this does not blow the 1e-5
    non-orthonormality of a float32 matrix up to 1e-3 rad."""
    d = np.linalg.norm(np.asarray(Ra, np.float64) - np.asarray(Rb, np.float64))
    return float(2 * np.arcsin(min(1.0, d / (2 * np.sqrt(2)))))
