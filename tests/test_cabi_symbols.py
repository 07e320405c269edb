"""The C-ABI library loads on a CPU-only box and exports every symbol the header declares
(no compute calls here).  Host-only entry points are exercised."""
import ctypes as C

import pytest

from b200 import _native as N
from b200 import atlas as A
from csrc_build import ensure_built


@pytest.fixture(scope="module", autouse=True)
def _built():
    ensure_built()


def test_every_declared_symbol_is_exported_and_bound():
    declared = set(N.header_functions())
    assert declared, "header parse found nothing"
    assert declared == set(N.SIGNATURES), declared ^ set(N.SIGNATURES)
    handle = C.CDLL(N.LIB_PATH)
    for name in declared:
        assert hasattr(handle, name), name
    assert N.lib().b200_version() >= 100


def test_flat_layout_matches_reference_parameter_counts():
    lib = N.lib()
    m = A.make_desc(**A.MAPPING_DESC)
    a = A.make_desc(**A.ATLAS_DESC)
    mw, mb, mt = A.mlp_layout(m)
    aw, ab, at = A.mlp_layout(a)
    # unpadded counts printed by IMLP.__init__ in the reference: 264706 / 416379
    assert sum(k * n + n for k, n in A.layer_dims(m)) == 264706
    assert sum(k * n + n for k, n in A.layer_dims(a)) == 416379
    assert mt == 264708 and at == 416380 and lib.b200_atlas_param_floats() == mt + at
    assert all(o % 4 == 0 for o in mw + mb + aw + ab)
    assert A.layer_dims(a)[4] == (296, 256) and A.layer_dims(a)[7] == (296, 3)


def test_invalid_descriptor_is_rejected_with_message():
    bad = N.MlpDesc(3, 2, 256, 1, 0, 0, 1, 0)          # a 1-layer network is not an IMLP
    assert N.lib().b200_mlp_layout(C.byref(bad), None, None) == -1
    assert b"invalid" in N.lib().b200_last_error()


def test_workspace_sizes_are_positive_and_monotone():
    lib = N.lib()
    small = N.AtlasConfig(1000, 1, N.PREC_FP32, 768, 0.8, 1, 100, 5000, 1000, 1, 5, 500)
    big = N.AtlasConfig(10000, 1, N.PREC_FP32, 768, 0.8, 1, 100, 5000, 1000, 1, 5, 500)
    a, b = lib.b200_atlas_workspace_bytes(C.byref(small)), lib.b200_atlas_workspace_bytes(C.byref(big))
    assert 0 < a < b
    assert lib.b200_render_workspace_bytes(1000) > 0
    assert lib.b200_render_workspace_bytes(0) == -1


def test_frame_ranges_partition_the_video():
    for T in (80, 7, 16):
        for world in (1, 2, 4, 8):
            spans = [A.frame_range(r, world, T) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == T
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            assert max(b - a for a, b in spans) - min(b - a for a, b in spans) <= 1
