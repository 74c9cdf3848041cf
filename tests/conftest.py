import json
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

GOLDEN = Path(__file__).resolve().parent / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    import torch

    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


def fnv1a64_words(words) -> int:
    f = 1469598103934665603
    for x in np.asarray(words, dtype=np.uint32).tolist():
        f = ((f ^ x) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
    return f


@pytest.fixture(scope="session")
def roc_golden():
    """Golden vectors produced by the unmodified reference (tests/golden/gen_golden.py)."""
    z = np.load(GOLDEN / "roc_golden.npz")
    meta = json.loads(str(z["meta"]))
    cases = []
    for i, (tag, p, head, final_head, final_nwords) in enumerate(meta):
        cases.append(
            dict(tag=tag, p=int(p), head=int(head), final_head=int(final_head), final_nwords=int(final_nwords),
                 ids=z[f"ids_{i}"], words=z[f"words_{i}"], order=z[f"order_{i}"], dec=z[f"dec_{i}"])
        )
    return cases


@pytest.fixture(scope="session")
def roc_kat():
    return json.load(open(GOLDEN / "roc_kat.json"))


def codec_workload_ids(seed: int, n: int = 65000, nbits: int = 20) -> np.ndarray:
    """Input stream of the reference's test_codec.cpp:62-82 (mt19937(seed) & mask, repeats skipped)."""
    raw = np.random.RandomState(seed)._bit_generator.random_raw(4 * n).astype(np.uint64) & ((1 << nbits) - 1)
    _, first = np.unique(raw, return_index=True)
    first.sort()
    return raw[first][:n]
