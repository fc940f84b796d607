"""GPU test that pins the oracle: the reference's OWN kernels (oracle/_ref, compiled unmodified
from /root/reference/cuda_lib/src/image_filters.cu for sm_100a) run on this GPU and must equal
the CPU oracle byte for byte, level by level."""
import numpy as np
import pytest

from oracle import oracle as O
from tests import synth

pytestmark = pytest.mark.gpu

ENUM = {"gaussian": {1: 1, 2: 3}, "box": {1: 1, 2: 2}, "sobel": {1: 1, 2: 2}}


def _ref(kind, img, level, sigma=2.0, radius=3):
    import torch
    d_in = torch.from_numpy(img).cuda()
    d_out = torch.empty_like(d_in)
    torch.cuda.synchronize()
    h, w, c = img.shape
    rc, _ = O.ref_call(kind, d_in.data_ptr(), d_out.data_ptr(), w, h, c, ENUM[kind][level], sigma, radius)
    torch.cuda.synchronize()
    assert rc == 0
    return d_out.cpu().numpy()


@pytest.mark.skipif(not O.ref_available(), reason="oracle/_ref not built")
@pytest.mark.parametrize("c", [1, 3, 4])
@pytest.mark.parametrize("kind", ["uniform", "smooth"])
def test_oracle_equals_reference_kernels(c, kind):
    img = synth.KINDS[kind](97, 131, c, seed=11 + c)
    for level in (1, 2):
        for r, s in ((1, 0.7), (3, 2.0), (9, 4.0)):
            assert np.array_equal(_ref("gaussian", img, level, s, r), O.gaussian_blur(img, s, r)), ("gaussian", level, r)
        for r in (1, 3, 5, 16):
            assert np.array_equal(_ref("box", img, level, radius=r), O.box_blur(img, r)), ("box", level, r)
        assert np.array_equal(_ref("sobel", img, level), O.sobel(img, level)), ("sobel", level)
    # level 1 has no radius limit
    assert np.array_equal(_ref("box", img, 1, radius=31), O.box_blur(img, 31))
    assert np.array_equal(_ref("gaussian", img, 1, 10.0, 31), O.gaussian_blur(img, 10.0, 31))


@pytest.mark.skipif(not O.ref_available(), reason="oracle/_ref not built")
def test_reference_rejects_shared_level_for_gaussian():
    """image_filters.cu:693-696: SHARED_MEMORY is not implemented for Gaussian."""
    import torch
    x = torch.zeros(8, 8, 3, dtype=torch.uint8, device="cuda")
    y = torch.empty_like(x)
    rc, _ = O.ref_call("gaussian", x.data_ptr(), y.data_ptr(), 8, 8, 3, 2, 2.0, 3)
    assert rc == 801
