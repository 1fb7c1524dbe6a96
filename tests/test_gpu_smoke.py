import pytest

from tests import fixtures as fx

pytestmark = pytest.mark.gpu


@pytest.mark.skipif(not fx.have_models(), reason="engines not built")
def test_smoke_detect_plus_locate():
    fx.run_smoke(verbose=True)
