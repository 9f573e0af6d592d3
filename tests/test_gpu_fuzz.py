"""A short run of the randomised differential test (tools/fuzz_parity.py): random small bilayers,
every kernel family against the oracle port."""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))


@pytest.mark.parametrize("seed", [11, 12])
def test_random_bilayers_match_the_port(seed):
    import fuzz_parity
    rng = np.random.default_rng(seed)
    for k in range(12):
        desc, bad = fuzz_parity.one_case(rng, k)
        assert not bad, (desc, bad)
