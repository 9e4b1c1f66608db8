"""Camera::UndistortImage (cv::remap, src/camera.cc:161-182) on the device, fused into the resize kernel.
Bit-exact: integer pixel arithmetic.  The oracle's cv_remap_u8 is itself pinned against cv2.remap on the CPU (tests/test_host_logic.py)."""
import numpy as np
import pytest

import _parity as P

pytestmark = pytest.mark.gpu

W, H = 752, 480


def _maps():
    from oracle import host
    # EuRoC-like radial-tangential cameras (configs/camera/euroc.yaml shape): left and right differ
    l = host.radtan_rectify_maps(W, H, 458.654, 457.296, 367.215, 248.375, -0.28340811, 0.07395907, 0.00019359, 1.76187114e-05, new_cx=362.0, new_cy=250.0, new_f_scale=0.78)
    r = host.radtan_rectify_maps(W, H, 457.587, 456.134, 379.999, 255.238, -0.28368365, 0.07451284, -0.00010473, -3.55590700e-05, new_cx=362.0, new_cy=250.0, new_f_scale=0.78)
    return l, r


def test_undistort_and_fused_resize_bit_exact():
    from airslam_b200 import capi
    from oracle import host, synth
    raw_l, raw_r, _ = synth.stereo_pair(W, H, 301)
    (lx, ly), (rx, ry) = _maps()
    rect_l, rect_r = host.cv_remap_u8(raw_l, lx, ly), host.cv_remap_u8(raw_r, rx, ry)
    assert (rect_l == 0).sum() > 100          # the maps leave the raw image at the borders: BORDER_CONSTANT(0) is exercised
    ctx = capi.Context(max_batch=1, enable_plnet=0)
    try:
        ctx.set_rectify_maps(0, lx, ly)
        ctx.set_rectify_maps(1, rx, ry)
        P.exact("remap: airfe_undistort == cv::remap restatement (left)", np.array_equal(ctx.undistort(0, raw_l), rect_l))
        P.exact("remap: airfe_undistort == cv::remap restatement (right)", np.array_equal(ctx.undistort(1, raw_r), rect_r))
        # rectification off, CPU-rectified frames in  ==  rectification on, raw frames in
        ref = ctx.stereo_batch(capi.NET_SUPERPOINT, capi.MATCHER_LIGHTGLUE, rect_l[None], rect_r[None])[0]
        ref_mono = ctx.detect_batch(capi.NET_SUPERPOINT, rect_l[None])[0][0]
        ctx.set_rectify(2)
        got = ctx.stereo_batch(capi.NET_SUPERPOINT, capi.MATCHER_LIGHTGLUE, raw_l[None], raw_r[None])[0]
        for i, rect in ((0, rect_l), (1, rect_r)):
            x16 = ctx.debug_read(capi.NET_SUPERPOINT, "x16", i, np.float16, (512, 512))
            P.exact("remap fused into resize: network input == resize(remap(raw)) (%s)" % ("left", "right")[i], np.array_equal(x16, host.process_image(rect)[0, 0].astype(np.float16)))
        P.exact("stereo entry on raw frames == on CPU-rectified frames (features, matches)",
                np.array_equal(got["feat_l"], ref["feat_l"]) and np.array_equal(got["feat_r"], ref["feat_r"]) and np.array_equal(got["matches"][0], ref["matches"][0]) and
                np.array_equal(got["matches"][1], ref["matches"][1]))
        ctx.set_rectify(1)
        P.exact("mono detect on a raw frame == on the CPU-rectified frame", np.array_equal(ctx.detect_batch(capi.NET_SUPERPOINT, raw_l[None])[0][0], ref_mono))
        ctx.set_rectify(0)
        assert np.array_equal(ctx.detect_batch(capi.NET_SUPERPOINT, rect_l[None])[0][0], ref_mono)
        # wrong frame size with rectification on is an error, not a silent pass-through
        ctx.set_rectify(1)
        with pytest.raises(capi.AirfeError):
            ctx.detect_batch(capi.NET_SUPERPOINT, np.zeros((1, 240, 376), np.uint8))
    finally:
        ctx.close()
