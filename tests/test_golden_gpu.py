"""The CUDA blocks against the golden outputs of the reference's own functions."""
import pytest

import golden_cases as G

pytestmark = pytest.mark.gpu

CASES = [(f, c, a) for f, c, a in G.all_cases() if c["family"] not in ("percentile", "labelled")]


@pytest.mark.parametrize("family,case,arrays", CASES,
                         ids=["{}-{}-{}".format(f, c["family"], c["id"]) for f, c, _ in CASES])
def test_cuda_matches_reference(family, case, arrays):
    G.compare(case, G.run_product(case, arrays), G.expected_of(case, arrays))
