"""The shim with a real SIZE layout: lelt = nelt + 3 (tests/3dboxper/SIZE:15 pads lelt by 3), so
every vector field in COMMON is (lpts1,3) with lpts1 > npts and a .usr's ADE arrays are (lpts,k).
b200_copy_all_in / b200_update_device / b200_update_host must honour those leading dimensions
(nekcem_b200_set_array_ld / get_array_ld / set_leading_dims).  The other drop-in tests run with
lelt = nelt, where the leading dimension equals npts.

NOT YET RUN ON HARDWARE: written after this round's GPU budget was exhausted (the host-side
packing is covered on CPU by tests/test_abi.py and tests/test_reference_pin.py::
test_pin_padded_size_layout). Green on B200 since the round-1 driver run (GPUTEST_r01: xpassed)."""
import os
import subprocess
import sys

import pytest

pytestmark = [pytest.mark.gpu]


ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("which", ["pml", "drude"])
def test_dropin_padded_size(which):
    """pml: 3ddielectric geometry with PML (pmlsigma, pmlbn, pmldn are (lpts1,3) too); drude:
    tests/drude with jn(lpts,3), params(lpts,2) dimensioned by the padded SIZE and registered
    through the .usr's usersrc -> cem_maxwell_drude twin.  Run in a process of its own
    (tests/drivers/padded_dropin_check.py): the twins exit(1) on a library error."""
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "drivers", "padded_dropin_check.py"), which],
                       capture_output=True, text=True, timeout=600)
    if r.returncode == 77:
        pytest.skip(r.stdout.strip())
    assert r.returncode == 0, r.stdout + r.stderr
