"""GPU parity of the whole model (embedding, encoders, classifier, fused decoder, postnet, loss) against the golden
vectors recorded from the unmodified reference, including every parameter gradient and the BN running statistics."""
import pytest
import torch

import model_cases
from helpers import GOLDEN_CASES

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module', autouse=True)
def _built():
    import __graft_entry__ as entry
    entry.build()
    assert torch.cuda.is_available()


@pytest.mark.parametrize('name', GOLDEN_CASES)
def test_model_matches_reference_golden(name):
    model_cases.run_golden(name, check_grads=True, verbose=True)
