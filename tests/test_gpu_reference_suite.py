"""The reference's own tests (tests/reference_suite.py) against the CUDA library, through the C ABI."""
import pytest

import reference_suite as RS

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("check", RS.ALL_CHECKS, ids=lambda c: c.__name__)
def test_device_reference_suite(ifb, device, check):
    device.reset_launch_count()
    check(ifb, None)
    assert device.launch_count() > 0, "no CUDA kernel was launched: the product path did not run"
