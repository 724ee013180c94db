"""oracle/f2c_lite.py (TEST INFRASTRUCTURE: the translator through which the reference's own
Fortran hot path is executed here) on a page of Fortran whose results are fixed by the language:
integer division and mixed-mode promotion, ** precedence, mod / int / nint / sign, COMPLEX
arithmetic with Fortran's promotion of real operands (the sign of zero decides the branch of
sqrt), labelled and block DO loops (zero trip, negative stride, exit / cycle, the loop variable
after completion), column-major adjustable arrays, SAVE + DATA, implicit typing.  The pin of the
oracle against the translated reference (tests/test_reference_pin.py) rests on these."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


@pytest.fixture(scope="module")
def lib(tmp_path_factory):
    from oracle import f2c_lite     # (package import: never put oracle/ itself on sys.path --
    #                                 `import oracle` would then find oracle/oracle.py)
    d = tmp_path_factory.mktemp("f2c")
    text, _ = f2c_lite.translate([(os.path.join(HERE, "f2c_semantics.F"),
                                   ["t_arith", "t_loops", "t_count", "t_implicit"])], [HERE], ())
    src, so = str(d / "gen.c"), str(d / "gen.so")
    open(src, "w").write(text)
    subprocess.check_call(["gcc", "-O2", "-std=gnu11", "-ffp-contract=off", "-fcx-fortran-rules",
                           "-fPIC", "-shared", "-w", "-o", so, src, "-lm"])
    return C.CDLL(so)


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def test_arithmetic_semantics(lib):
    r = np.zeros(20)
    lib.t_arith_(_dp(r))
    want = [3, -3, 6, 7, -1, 49, 9, -9, 9, 0, 0.75, -2, -3, -3, 5, 2,
            2,   # aimag(sqrt((-4,0)))
            2,   # aimag(sqrt(1.0-(5,0))): the real 1.0 is promoted to (1,+0) first, so the
                 # imaginary part of the difference is +0 and sqrt takes the upper branch
            5, 5]
    assert np.array_equal(r, np.array(want, dtype=float)), r


def test_loop_and_array_semantics(lib):
    a, s = np.zeros(12), np.zeros(6)
    lib.t_loops_(C.byref(C.c_int(3)), C.byref(C.c_int(4)), _dp(a), _dp(s))
    assert np.array_equal(s, [12, 0, 22, 16, 4, 23]), s
    assert np.array_equal(a, [11, 21, 31, 12, 22, 32, 13, 23, 33, 14, 24, 34])   # column-major


def test_save_data_and_implicit_typing(lib):
    lib.t_count_.restype = C.c_double
    assert [lib.t_count_() for _ in range(3)] == [1.0, 2.0, 3.0]
    r = np.zeros(3)
    lib.t_implicit_(_dp(r))
    assert np.array_equal(r, [3.0, 3.5, 3.5])
