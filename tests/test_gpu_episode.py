"""GPU parity of dana_b200.episode (C ABI: dana_episode_resize) against oracle/episode_oracle.py, the numpy
restatement of the reference's loaders (blob.py:35-52, fs_loader.py:113-138, inference_loader.py:95-109) that
tests/test_oracle_pins.py pins to the real cv2.  Same float32 operation order on both sides: the bar is 1e-5 of the
8-bit pixel scale (2.6e-3 absolute would be one grey level; we ask for 2e-4), zero padding exact."""
import numpy as np
import pytest
import torch

import episode_oracle as E

pytestmark = pytest.mark.gpu
MEANS = [102.9801, 115.9465, 122.7717]


def _close(got, want):
    got = got.cpu().numpy()
    assert got.shape == want.shape
    np.testing.assert_allclose(got, want, rtol=0, atol=2e-4)


@pytest.mark.parametrize("h,w,target", [(375, 500, 600), (480, 640, 600), (333, 500, 600), (600, 901, 600), (50, 31, 48)])
@pytest.mark.parametrize("dtype", ["u8", "f32"])
def test_prep_im_for_blob(h, w, target, dtype):
    import dana_b200  # noqa: F401
    from dana_b200 import episode
    rs = np.random.RandomState(h + w)
    im = rs.randint(0, 256, size=(h, w, 3)).astype(np.uint8)
    src = im if dtype == "u8" else im.astype(np.float32)
    want, scale = E.prep_im_for_blob(src, MEANS, target)
    got, s2 = episode.prep_im_for_blob(torch.from_numpy(src).cuda(), MEANS, target, 1000)
    assert s2 == scale
    _close(got, np.ascontiguousarray(np.transpose(want, (2, 0, 1))))


def test_prep_into_padded_blob():
    """im_list_to_blob (blob.py:17-33): the image sits in the top-left corner of a larger zeroed canvas."""
    import dana_b200  # noqa: F401
    from dana_b200 import episode
    rs = np.random.RandomState(5)
    im = rs.randint(0, 256, size=(120, 200, 3)).astype(np.uint8)
    want, _ = E.prep_im_for_blob(im, MEANS, 150)
    canvas = torch.full((3, 160, 260), 7.0, device="cuda")
    episode.prep_im_for_blob(torch.from_numpy(im).cuda(), MEANS, 150, out=canvas)
    ref = np.zeros((3, 160, 260), dtype=np.float32)
    ref[:, :want.shape[0], :want.shape[1]] = np.transpose(want, (2, 0, 1))
    _close(canvas, ref)


@pytest.mark.parametrize("box,scale", [((10.2, 20.7, 60.9, 50.1), 1.5), ((0, 0, 139, 89), 1.0), ((100, 5, 139.9, 89.4), 1.0),
                                       ((30, 10, 45, 80), 2.0), ((5, 5, 6, 6), 1.0)])
def test_support_from_box(box, scale):
    import dana_b200  # noqa: F401
    from dana_b200 import episode
    rs = np.random.RandomState(int(box[0] * 10))
    im = (rs.standard_normal((int(90 * scale), int(140 * scale), 3)) * 50).astype(np.float32)
    want = E.support_from_box(im, box, scale, 320)
    got = episode.support_from_box(torch.from_numpy(im).cuda(), box, scale, 320)
    _close(got, want)


@pytest.mark.parametrize("h,w", [(200, 120), (120, 200), (320, 320), (17, 400)])
def test_support_from_image(h, w):
    import dana_b200  # noqa: F401
    from dana_b200 import episode
    rs = np.random.RandomState(h * 7 + w)
    im = rs.randint(0, 256, size=(h, w, 3)).astype(np.uint8)
    want = E.support_from_image(im, MEANS, 320)
    got = episode.support_from_image(torch.from_numpy(im).cuda(), MEANS, 320)
    _close(got, want)


def test_episode_rejects_bad_arguments():
    import dana_b200  # noqa: F401
    from dana_b200 import episode
    from dana_b200._lib import DanaError
    im = torch.zeros((8, 8, 3), dtype=torch.uint8, device="cuda")
    out = torch.empty((3, 4, 4), device="cuda")
    with pytest.raises(DanaError):
        episode.resize_into(im, (4, 4, 8, 8), (1.0, 1.0), (4, 4), out)       # crop leaves the image
    with pytest.raises(DanaError):
        episode.resize_into(im, (0, 0, 8, 8), (1.0, 1.0), (5, 5), out)       # resized extent exceeds the canvas
    with pytest.raises(Exception):
        episode.prep_im_for_blob(torch.zeros((8, 8, 3), dtype=torch.uint8), MEANS, 8)   # CPU tensor: no fallback
