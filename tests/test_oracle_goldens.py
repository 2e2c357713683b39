"""Pins the CPU oracle (oracle/oracle.cpp) against the reference's own goldens and identities."""
import numpy as np
import pytest

import reference_suite as RS


@pytest.mark.parametrize("check", RS.ALL_CHECKS, ids=lambda c: c.__name__)
def test_oracle_reference_suite(ifb, oracle, check):
    check(ifb, oracle)


def test_oracle_lemire_matches_separable_form(ifb, oracle):
    """The oracle's min/max is the truncated-window ground truth; check it against a restatement of
    the reference's actual streaming algorithm (src/mapwindow.jl:426-473), incl. even windows."""
    import ctypes as C
    fn = oracle.dll.b2f_oracle_lemire_axis0
    fn.argtypes = [C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_int64, C.c_int64, C.c_int64]
    rng = np.random.default_rng(0)
    for n, m in ((5, 1), (17, 3), (64, 5), (9, 2)):
        for w in (2, 3, 4, 5, 7, 8, 9):
            if (w >> 1) >= n:
                continue
            A = np.asfortranarray(rng.integers(0, 50, size=(n, m)).astype(np.float64))
            mn, mx = A.copy(order="F"), A.copy(order="F")
            assert fn(mn.ctypes.data_as(C.POINTER(C.c_double)), mx.ctypes.data_as(C.POINTER(C.c_double)), n, m, w) == 0
            mm = ifb.mapwindow(ifb.extrema, A, (w, 1), _library=oracle)
            assert np.array_equal(mm["min"], mn), (n, m, w)
            assert np.array_equal(mm["max"], mx), (n, m, w)
