"""Horizontal bands of one canvas (second row): every spec scene that does not use a surface as a compositor
parameter is rendered as a stack of 32-row band surfaces -- the same calls replayed on each band -- and must equal the
oracle's full-canvas render exactly as the unbanded device render does."""
import numpy as np
import pytest

from tests import golden_util, specs
from tests.banded_backend import BandedBackend
from tests.test_gpu_scenes import FLOAT_SCENES

pytestmark = pytest.mark.gpu

CASES = [(stem, aa) for stem in sorted(specs.PATH_SCENES) for aa, _ in golden_util.AA_SUFFIX]


@pytest.mark.parametrize("stem,aa", CASES, ids=[f"{s}-{a.name}" for s, a in CASES])
def test_banded_render_equals_oracle(cuda, oracle, stem, aa):
    try:
        got = specs.PATH_SCENES[stem](specs.bind(BandedBackend(cuda, 32)), aa)
    except NotImplementedError as e:
        pytest.skip(str(e))
    ref = specs.PATH_SCENES[stem](specs.bind(oracle), aa)
    g, r = got.pixels().astype(np.int32), ref.pixels().astype(np.int32)
    assert g.shape == r.shape
    d = np.abs(g - r)
    if stem in FLOAT_SCENES:
        assert d.max() <= 1
    else:
        assert d.max() == 0, f"{int((d != 0).any(axis=-1).sum())} pixels differ"
