"""Clear / viewport / frame-boundary call sequences (tests/sequences.py): the CPU restatement against the reference here, the
CUDA product against the restatement on the GPU box — bit for bit on every observable buffer."""
import os

import numpy as np
import pytest

from conftest import REF_LIB
from sequences import SEQUENCES


@pytest.mark.parametrize("name", list(SEQUENCES))
def test_sequence_oracle_equals_reference(oracle, reference, name):
    a, b = SEQUENCES[name](oracle), SEQUENCES[name](reference)
    assert len(a) == len(b)
    for i, (x, y) in enumerate(zip(a, b)):
        assert np.array_equal(x, y), f"{name}: buffer {i}"


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(SEQUENCES))
def test_sequence_cuda_equals_oracle(cuda, oracle, name):
    a, b = SEQUENCES[name](cuda), SEQUENCES[name](oracle)
    assert len(a) == len(b)
    for i, (x, y) in enumerate(zip(a, b)):
        assert np.array_equal(x, y), f"{name}: buffer {i}: {(x != y).sum()} values differ"
