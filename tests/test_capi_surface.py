"""CPU tests of the drop-in boundary: the shared library loads, exports every symbol
include/mvs_b200.h declares, and refuses to compute without a GPU (no CPU fallback)."""
import numpy as np
import pytest

import multiviewstitch_b200 as m
from multiviewstitch_b200 import api


def test_library_exports_every_declared_symbol():
    L = m.lib()
    declared = m.declared_symbols()
    assert len(declared) > 40
    missing = [s for s in declared if not hasattr(L, s)]
    assert not missing, f"declared in include/mvs_b200.h but not exported: {missing}"


def test_no_cpu_fallback_without_device():
    if api.device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(api.MvsError) as e:
        api.Context(0)
    assert e.value.status == -3
    with pytest.raises(api.MvsError):
        api.nn_search(np.zeros((4, 3), np.float32), np.zeros((4, 3), np.float32), 1)


def test_draw_triples_matches_shuffle_semantics():
    # Common/Utils.h:25-34: 3 distinct indices, ascending, uniform over n
    tr = api.srt_draw_triples(10, 2000, 1)
    assert tr.min() >= 0 and tr.max() <= 9
    assert np.all(tr[:, 0] < tr[:, 1]) and np.all(tr[:, 1] < tr[:, 2])
    counts = np.bincount(tr.ravel(), minlength=10)
    assert counts.min() > 450 and counts.max() < 750
    import oracle
    assert np.array_equal(tr, oracle.draw_triples(10, 2000, 1))
