import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture(scope="session")
def model_npz():
    return np.load(os.path.join(GOLDEN, "model_dancedb.npz"))


@pytest.fixture(scope="session")
def pose_model():
    from dragposer_b200 import model

    return model.load_folded_npz(os.path.join(GOLDEN, "model_dancedb.npz"))


@pytest.fixture(scope="session")
def temporal_model():
    from dragposer_b200 import model

    return model.temporal_from_state(model.random_temporal_state(2222))


@pytest.fixture(scope="session")
def port_weights(model_npz):
    import dragposer_port as port

    return port.PortWeights(model_npz)


@pytest.fixture(scope="session")
def engine_factory(pose_model, temporal_model, model_npz):
    """Builds BatchedDragPose engines on cuda:0 (GPU tests only)."""
    from dragposer_b200.engine import BatchedDragPose

    made = []

    def make(max_clips):
        e = BatchedDragPose(pose_model, model_npz["offsets"], temporal_model, max_clips)
        made.append(e)
        return e

    yield make
    for e in made:
        e.close()
