import hashlib
import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLD = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")


def load_ref_mapped():
    with open(os.path.join(GOLD, "ref_mapped.json")) as f:
        return json.load(f)


def load_lbf_index():
    with open(os.path.join(GOLD, "lbf", "index.json")) as f:
        return json.load(f)


def unpack_outputs(entry, batch=1000):
    return {k: np.unpackbits(np.frombuffer(bytes.fromhex(v), dtype=np.uint8))[:batch] for k, v in entry["outputs"].items()}


def out_hash(outputs):
    """sha256 over str(name) bytes + uint8 array per output in order (SURVEY.md Appendix F / oracle/gen_golden.py)."""
    h = hashlib.sha256()
    for name, arr in outputs.items():
        h.update(str(name).encode())
        h.update(np.asarray(arr).astype(np.uint8).tobytes())
    return h.hexdigest()


def selfcheck_inputs(names, batch=1000, seed=42):
    """reference map_circuit.py:137-139"""
    np.random.seed(seed)
    return {nm: np.random.randint(0, 2, (batch)) for nm in names}


@pytest.fixture(scope="session")
def ref_mapped():
    return load_ref_mapped()


@pytest.fixture(scope="session")
def lbf_index():
    return load_lbf_index()


def read_golden_lbf(fn):
    from tfhe_fbs_map_b200.formats import read_lbf, read_lbf_file
    if fn.endswith(".gz"):
        import gzip
        with gzip.open(os.path.join(GOLD, "lbf", fn), "rt") as f:
            return read_lbf(f.read())
    return read_lbf_file(os.path.join(GOLD, "lbf", fn))


def read_golden_blif(name):
    from tfhe_fbs_map_b200.formats import parse_blif_file
    return parse_blif_file(os.path.join(GOLD, "blif", f"{name}.blif"))


def have_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False
