"""GPU tests of the row-band split: one frame cut into row bands over several GPUs, halos / lattice / band rows moved
with peer-to-peer copies, result bit-identical to the single-device path and to the reference.  With one GPU the same
device is listed several times (all the exchange logic runs, the "peer" copies are local); with two or more GPUs the
bands live on different devices (NVLink P2P)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def device_list(svb, n):
    nd = svb.device_count()
    return [i % nd for i in range(n)]


@pytest.mark.parametrize("n_bands", [1, 2, 3, 4])
def test_band_split_matches_reference_kitti(svb, kitti_gray, golden, n_bands):
    L, R = kitti_gray["L0"], kitti_gray["R0"]
    H, W = L.shape
    g = svb.BandGroup(svb.default_params(svb.PIPELINE), W, H, device_list(svb, n_bands))
    try:
        D1, D2 = g.process(L, R)
        assert np.array_equal(D1, golden["pipeline_0_D1"])
        st = g.stats()
        assert st["bands"] == n_bands and st["support_points"] == len(golden["pipeline_0_support"]) and st["gpu_ms"] > 0
        if n_bands > 1 and svb.device_count() > 1:
            # distinct devices: the bands really live in different GPUs' memories and the copies cross NVLink
            assert len(set(device_list(svb, n_bands))) == min(n_bands, svb.device_count()) and st["peer_links"] > 0
        if n_bands > 1:
            # halos: 2 images x 2 rows x 16 W bytes per band edge and side; lattice rows; both maps' band rows
            assert st["p2p_copies"] >= 4 * (n_bands - 1) + (n_bands - 1) + 2 * (n_bands - 1)
            assert st["p2p_bytes"] > 2 * 4 * W * (H // n_bands) * (n_bands - 1) // 2
        # robotics preset (no corners, different filters) through the same group size
    finally:
        g.close()
    g = svb.BandGroup(svb.default_params(svb.ROBOTICS), W, H, device_list(svb, n_bands))
    try:
        D1, _ = g.process(kitti_gray["L7"], kitti_gray["R7"])
        assert np.array_equal(D1, golden["robotics_7_D1"])
    finally:
        g.close()


def test_band_split_4k_disp512(svb, ref):
    """BASELINE.json configs[3]: 3840x2160, disparity range 512, L/R check and all post-filters, bands over the GPUs."""
    W, H = 3840, 2160
    L, R = svb.synth_pair(9, W, H, 0)
    p = svb.default_params(svb.MIDDLEBURY, disp_max=511)
    nd = max(2, min(svb.device_count(), 8))
    g = svb.BandGroup(p, W, H, device_list(svb, nd))
    try:
        D1, D2 = g.process(L, R)
        R1, R2, _ = ref.process(ref.params(1, disp_max=511), L, R)
        assert np.array_equal(D1, R1) and np.array_equal(D2, R2)
        st = g.stats()
        assert st["p2p_bytes"] >= 2 * 4 * W * H * (nd - 1) // nd
    finally:
        g.close()


def test_band_split_few_support_points(svb):
    W, H = 640, 240
    flat = np.full((H, W), 100, np.uint8)
    g = svb.BandGroup(svb.default_params(svb.ROBOTICS), W, H, device_list(svb, 2))
    try:
        with pytest.raises(svb.SvbError) as e:
            g.process(flat, flat)
        assert e.value.code == svb.ERR_FEW_SUPPORT
    finally:
        g.close()
