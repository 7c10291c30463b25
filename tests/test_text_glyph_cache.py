"""Text runs through the glyph cache (z2d_glyph_cache_add / z2d_fill_glyphs): the reference's three text scenes
(spec/074, 080 -- BASELINE config 1 -- and 085) rendered with the outlines resident in the backend and one transformation per
glyph from the host must equal, byte for byte, the renders from host-built node lists (which equal the goldens,
tests/test_oracle_goldens.py / tests/test_gpu_scenes.py)."""
import numpy as np
import pytest

from tests import specs
from tests.specs import ttf
from z2d_b200.abi import AntiAliasMode

TEXT_SCENES = [s for s in specs.PATH_SCENES if s.split("_")[0] in ("074", "080", "085")]
BACKENDS = ["oracle", pytest.param("cuda", marks=pytest.mark.gpu)]


def test_the_three_text_scenes_are_covered():
    assert len(TEXT_SCENES) == 3


@pytest.mark.parametrize("backend", BACKENDS)
@pytest.mark.parametrize("aa", [AntiAliasMode.default, AntiAliasMode.none, AntiAliasMode.supersample_4x], ids=lambda a: a.name)
@pytest.mark.parametrize("stem", TEXT_SCENES)
def test_cached_text_equals_host_built_nodes(request, backend, stem, aa):
    be = request.getfixturevalue(backend)
    plain = specs.PATH_SCENES[stem](specs.bind(be), aa).download().copy()
    ttf.use_glyph_cache(True)
    try:
        cached = specs.PATH_SCENES[stem](specs.bind(be), aa).download().copy()
    finally:
        ttf.use_glyph_cache(False)
    assert np.array_equal(plain, cached), f"{int((plain != cached).sum())} bytes differ"
    assert plain.any()
