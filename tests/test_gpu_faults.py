"""Failure detection: a rank that never signals must trip the in-kernel watchdog of its neighbours
(status flag + exception), not hang the GPU. The reference's only failure handling is MPI_Abort /
exit() on bad arguments."""
import pytest

from conftest import random_image
from parallel_convolution_b200.models.filters import Filter

pytestmark = pytest.mark.gpu


def test_dropped_signal_trips_watchdog(rng):
    from parallel_convolution_b200.parallel.decomposition import TileGrid
    from parallel_convolution_b200.parallel.fused import VirtualCluster
    img = random_image(rng, 128, 256, 1)
    g = TileGrid.create(128, 256, 1, 2)
    vc = VirtualCluster(g, Filter.builtin("gaussian"), halo_depth=2, mode="stream", timeout_ms=200.0)
    try:
        vc.upload(img)
        vc.rt.drivers[0].inject_drop_signal = True      # rank 0 computes but never tells rank 1
        vc.run(6)                                       # 3 launches; rank 1's 2nd launch must time out
        vc.download()
        st = vc.collect_status()
        assert st.watchdog, "the watchdog should have fired on the rank waiting for rank 0"
        assert st.max_wait_us >= 150_000
    finally:
        vc.close()


def test_healthy_run_reports_no_watchdog(rng):
    from parallel_convolution_b200 import ConvEngine
    img = random_image(rng, 128, 256, 3)
    with ConvEngine(width=256, height=128, channels=3, backend="fused", workers=4) as eng:
        eng.load_image(img)
        rep = eng.iterate(9)
        assert not rep.watchdog and rep.launches > 0 and rep.elapsed_ms > 0
