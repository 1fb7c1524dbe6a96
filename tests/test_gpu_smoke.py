import pytest

from tests import fixtures as fx

pytestmark = pytest.mark.gpu


@pytest.mark.skipif(not fx.have_models(), reason="engines not built")
def test_smoke_detect_plus_locate():
    fx.run_smoke(verbose=True)


def test_comm_roundtrip_single_rank():
    """The in-library exchange (rmr_comm_*: pack -> H2D -> ncclAllGather -> D2H on its own stream) with a world of one."""
    import numpy as np
    import rm_radar_b200 as rr
    from rm_radar_b200 import _lib
    comm = rr.Comm(rr.Comm.unique_id(), 0, 1, device=0, max_robots=8)
    recs = (_lib.RobotRec * 8)()
    recs[0].is_detected = 1; recs[0].label = 5; recs[0].confidence = 0.75; recs[0].is_located = 1
    recs[0].location[0] = 1.0; recs[0].location[1] = 2.0; recs[0].location[2] = 3.0
    recs[0].rect[2] = 4.0; recs[0].rect[3] = 5.0
    for _ in range(3):
        comm.publish(recs, 1)
        got = comm.collect()
    assert got.shape == (1, 8, 8)
    assert np.array_equal(got[0], rr.Comm.pack(recs, 1, 8))
    assert got[0, 0].tolist() == [1.0, 5.0, 0.75, 1.0, 1.0, 2.0, 3.0, 20.0]
