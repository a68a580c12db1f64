import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def cfg():
    from spark_tts_b200.config import BiCodecConfig
    return BiCodecConfig()


@pytest.fixture(scope="session")
def state_dict(cfg):
    """Synthetic checkpoint, seed 0 -- the weights every golden fixture was generated with."""
    from spark_tts_b200.synthetic import synthetic_state_dict
    return synthetic_state_dict(cfg, seed=0)


def golden_cases():
    import glob
    d = os.path.join(ROOT, "tests", "golden")
    return sorted(glob.glob(os.path.join(d, "detok_*.npz")))


def tokenize_golden_cases():
    import glob
    d = os.path.join(ROOT, "tests", "golden")
    return sorted(glob.glob(os.path.join(d, "tokenize_*.npz")))


@pytest.fixture(scope="session")
def state_dict_with_encoder(cfg, state_dict):
    """The synthetic checkpoint plus the encode-side tensors of the semantic tokenize row."""
    from spark_tts_b200.synthetic import synthetic_encoder_state_dict
    return {**state_dict, **synthetic_encoder_state_dict(cfg, seed=0)}


def speaker_golden_cases():
    import glob
    d = os.path.join(ROOT, "tests", "golden")
    return sorted(glob.glob(os.path.join(d, "speaker_*.npz")))


@pytest.fixture(scope="session")
def state_dict_with_speaker(cfg, state_dict):
    """The synthetic checkpoint plus every encode-side tensor (feature encoder + ECAPA / perceiver / FSQ project_in)."""
    from spark_tts_b200.synthetic import synthetic_encoder_state_dict, synthetic_speaker_state_dict
    return {**state_dict, **synthetic_encoder_state_dict(cfg, seed=0), **synthetic_speaker_state_dict(cfg, seed=0)}
