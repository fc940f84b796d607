"""GPU parity tests: the CUDA library, called through the C ABI (ctypes), against the oracle on
the same seeded inputs.

Bar: box blur bit-exact; Gaussian and Sobel within 1 LSB per channel.
TOL below is that stated tolerance; the kernels are written to reproduce the reference's
float32 operation order, so the tests additionally require exact equality (EXACT = True)."""
import ctypes

import numpy as np
import pytest

from gpu_image_processing_b200 import _lib, gpu_filters
from oracle import oracle as O
from tests import synth

pytestmark = pytest.mark.gpu

TOL = {"gaussian": 1, "box": 0, "sobel": 1}
EXACT = True

SHAPES = [(1, 1), (1, 2), (2, 1), (1, 37), (37, 1), (2, 2), (3, 3), (5, 4), (17, 33), (33, 17),
          (64, 64), (100, 300), (129, 1025), (270, 481)]


def _check(kind, got, want, what=""):
    assert got.shape == want.shape and got.dtype == np.uint8
    d = np.abs(got.astype(np.int16) - want.astype(np.int16))
    assert d.max() <= TOL[kind], (what, int(d.max()))
    if EXACT:
        assert np.array_equal(got, want), (what, "mismatching bytes", int((d > 0).sum()))


@pytest.fixture(params=[0, 1], ids=["fast", "general"])
def path(request):
    old = _lib.load().gip_set_path(request.param)
    yield request.param
    _lib.load().gip_set_path(old)


@pytest.mark.parametrize("c", [1, 3, 4])
@pytest.mark.parametrize("hw", SHAPES)
def test_box_blur_bit_exact(hw, c, path):
    img = synth.uniform(hw[0], hw[1], c, seed=hw[0] * 31 + hw[1] + c)
    for r in (0, 1, 2, 3, 5, 8, 16, 31):
        for level in (1, 2):
            out = gpu_filters.box_blur(img, radius=r, level=level)["image"]
            _check("box", out, O.box_blur(img, r), f"box {hw} c={c} r={r} L{level}")


@pytest.mark.parametrize("c", [1, 3, 4])
@pytest.mark.parametrize("hw", SHAPES)
def test_gaussian_blur(hw, c, path):
    img = synth.uniform(hw[0], hw[1], c, seed=hw[0] * 17 + hw[1] + c)
    for r, s in ((0, 1.0), (1, 0.5), (2, 1.0), (3, 2.0), (5, 2.5), (7, 3.0), (15, 5.0), (31, 10.0)):
        for level in (1, 2):
            out = gpu_filters.gaussian_blur(img, sigma=s, radius=r, level=level)["image"]
            _check("gaussian", out, O.gaussian_blur(img, s, r), f"gaussian {hw} c={c} r={r} L{level}")


@pytest.mark.parametrize("c", [1, 3, 4])
@pytest.mark.parametrize("hw", SHAPES)
@pytest.mark.parametrize("kind", ["uniform", "smooth"])
def test_sobel(hw, c, kind, path):
    img = synth.KINDS[kind](hw[0], hw[1], c, seed=hw[0] * 13 + hw[1] + c)
    for level in (1, 2):
        out = gpu_filters.sobel_edge_detection(img, level=level)["image"]
        _check("sobel", out, O.sobel(img, level), f"sobel {hw} c={c} L{level}")


def test_wide_radius_takes_the_general_path():
    img = synth.uniform(90, 140, 3, seed=5)
    _check("box", gpu_filters.box_blur(img, radius=40)["image"], O.box_blur(img, 40))
    _check("gaussian", gpu_filters.gaussian_blur(img, sigma=12.0, radius=40)["image"], O.gaussian_blur(img, 12.0, 40))


def test_special_images(path):
    for v in (0, 255):
        img = synth.constant(40, 70, 3, v)
        assert np.array_equal(gpu_filters.gaussian_blur(img)["image"], img)
        assert np.array_equal(gpu_filters.box_blur(img)["image"], img)
        assert not gpu_filters.sobel_edge_detection(img)["image"].any()
    img = synth.white_square(108, 192, 1)        # tests/test_gaussian_blur.cu:22-36
    _check("gaussian", gpu_filters.gaussian_blur(img, 2.0, 3)["image"], O.gaussian_blur(img, 2.0, 3))


def test_metrics_and_result_dict():
    img = synth.uniform(256, 256, 3)
    res = gpu_filters.gaussian_blur(img, sigma=2.0, radius=3, level=2)
    assert set(res) == {"image", "time_ms", "bandwidth_gbps", "fps"}
    assert res["time_ms"] > 0 and res["fps"] == pytest.approx(1000.0 / res["time_ms"], rel=1e-3)
    want_bw = 4 * img.size / (res["time_ms"] / 1000.0) / 2 ** 30          # image_filters.cu:905-906
    assert res["bandwidth_gbps"] == pytest.approx(want_bw, rel=1e-3)
    res = gpu_filters.sobel_edge_detection(img)
    want_bw = 2 * img.size / (res["time_ms"] / 1000.0) / 2 ** 30          # image_filters.cu:1711-1712
    assert res["bandwidth_gbps"] == pytest.approx(want_bw, rel=1e-3)


def test_dtype_cast_and_noncontiguous_input():
    img = synth.uniform(50, 60, 3)
    a = gpu_filters.box_blur(img.astype(np.int32), radius=2)["image"]
    b = gpu_filters.box_blur(img[:, ::-1][:, ::-1], radius=2)["image"]
    want = O.box_blur(img, 2)
    assert np.array_equal(a, want) and np.array_equal(b, want)


@pytest.mark.parametrize("c", [1, 3, 4])
def test_device_api_batched_and_cxx_entry_points(c, path):
    import torch
    from gpu_image_processing_b200 import device
    frames = np.stack([synth.uniform(72, 200, c, seed=s) for s in range(5)])
    x = torch.from_numpy(frames).cuda()
    g = device.gaussian_blur(x, 2.0, 3).cpu().numpy()
    b = device.box_blur(x, 4).cpu().numpy()
    s1 = device.sobel_edge_detection(x, 1).cpu().numpy()
    s2 = device.sobel_edge_detection(x, 2).cpu().numpy()
    for i in range(len(frames)):
        _check("gaussian", g[i], O.gaussian_blur(frames[i], 2.0, 3))
        _check("box", b[i], O.box_blur(frames[i], 4))
        _check("sobel", s1[i], O.sobel(frames[i], 1))
        _check("sobel", s2[i], O.sobel(frames[i], 2))
    # the reference's C++ symbols (image_filters.h:46-112) on raw device pointers
    L = ctypes.CDLL(_lib.LIB_PATH)
    fn = getattr(L, "_Z7boxBlurPhS_iiii17OptimizationLevelP18PerformanceMetrics")
    fn.argtypes = [ctypes.c_void_p, ctypes.c_void_p] + [ctypes.c_int] * 5 + [ctypes.c_void_p]
    m = _lib.Metrics()
    y = torch.empty_like(x[0])
    torch.cuda.synchronize()
    assert fn(x[0].data_ptr(), y.data_ptr(), 200, 72, c, 4, 1, ctypes.byref(m)) == 0
    assert np.array_equal(y.cpu().numpy(), b[0]) and m.time_ms > 0
    assert fn(x[0].data_ptr(), y.data_ptr(), 200, 72, c, 4, 3, ctypes.byref(m)) == 801   # bad level


@pytest.mark.parametrize("c", [1, 3, 4])
@pytest.mark.parametrize("nbands", [2, 3, 5])
def test_row_bands_stitch_to_the_whole_image(c, nbands, path):
    """N virtual ranks on one device; the stitched bands equal the whole image."""
    import torch
    L = _lib.load()
    h, w = 151, 233
    img = synth.uniform(h, w, c, seed=77 + c)
    x = torch.from_numpy(img).cuda()
    pitch = w * c
    cuts = [round(i * h / nbands) for i in range(nbands + 1)]
    stream = torch.cuda.current_stream().cuda_stream
    for kind, r in (("gaussian", 6), ("box", 9), ("sobel", 1)):
        out = torch.zeros_like(x)
        for i in range(nbands):
            y0, y1 = cuts[i], cuts[i + 1]
            ra, rb = min(r, y0), min(r, h - y1)
            band = x.data_ptr() + y0 * pitch
            above = x.data_ptr() + (y0 - ra) * pitch if ra else None
            below = x.data_ptr() + y1 * pitch if rb else None
            o = out.data_ptr() + y0 * pitch
            if kind == "gaussian":
                rc = L.gip_gaussian_blur_band(band, above, below, o, w, h, c, y0, y1 - y0, ra, rb, 3.0, r, 1, stream)
            elif kind == "box":
                rc = L.gip_box_blur_band(band, above, below, o, w, h, c, y0, y1 - y0, ra, rb, r, 1, stream)
            else:
                rc = L.gip_sobel_band(band, above, below, o, w, h, c, y0, y1 - y0, ra, rb, 1, stream)
            assert rc == 0
        torch.cuda.synchronize()
        want = {"gaussian": lambda: O.gaussian_blur(img, 3.0, r), "box": lambda: O.box_blur(img, r),
                "sobel": lambda: O.sobel(img, 1)}[kind]()
        _check(kind, out.cpu().numpy(), want, f"bands {kind} c={c} n={nbands}")
    # a band without its halo rows is refused
    assert L.gip_box_blur_band(x.data_ptr() + 50 * pitch, None, None, out.data_ptr(), w, h, c, 50, 20, 0, 0, 3, 1, stream) == 1


def _interior_band_check(kind, x_np, out_np, y0, y1, r, **kw):
    """Oracle on rows [y0-r, y1+r) of the input reproduces rows [y0, y1) of the full-image result."""
    h = x_np.shape[0]
    a, b = max(0, y0 - r), min(h, y1 + r)
    assert a == y0 - r and b == y1 + r
    sub = x_np[a:b]
    if kind == "gaussian":
        want = O.gaussian_blur(sub, kw["sigma"], r)
    elif kind == "box":
        want = O.box_blur(sub, r)
    else:
        want = O.sobel(sub, kw["level"])
    _check(kind, out_np[y0:y1], want[y0 - a:y0 - a + (y1 - y0)], f"{kind} interior rows {y0}:{y1}")


def test_baseline_config_shapes_against_oracle_bands():
    """BASELINE.json configs c1-c3 at full size: the fused kernels against the oracle on row bands
    (top edge, an interior band, bottom edge), and fast path == general path on the whole image."""
    import torch
    from gpu_image_processing_b200 import device
    L = _lib.load()
    cases = [("gaussian", (2146, 3239, 3), dict(sigma=2.0, radius=3)),
             ("box", (4096, 4096, 4), dict(radius=31)),
             ("box", (4096, 4096, 4), dict(radius=7)),
             ("sobel", (4320, 7680, 3), dict(level=1)),
             ("sobel", (4320, 7680, 3), dict(level=2))]
    for kind, (h, w, c), kw in cases:
        img = synth.uniform(h, w, c, seed=h + w)
        x = torch.from_numpy(img).cuda()
        run = {"gaussian": lambda: device.gaussian_blur(x, kw.get("sigma", 2.0), kw.get("radius", 3)),
               "box": lambda: device.box_blur(x, kw.get("radius", 3)),
               "sobel": lambda: device.sobel_edge_detection(x, kw.get("level", 1))}[kind]
        fast = run().cpu().numpy()
        old = L.gip_set_path(1)
        try:
            general = run().cpu().numpy()
        finally:
            L.gip_set_path(old)
        assert np.array_equal(fast, general), (kind, kw)
        r = kw.get("radius", 1)
        okw = dict(sigma=kw.get("sigma", 2.0), level=kw.get("level", 1))
        _interior_band_check(kind, img, fast, h // 2 - 40, h // 2 + 40, r, **okw)
        top = {"gaussian": lambda s: O.gaussian_blur(s, okw["sigma"], r), "box": lambda s: O.box_blur(s, r),
               "sobel": lambda s: O.sobel(s, okw["level"])}[kind]
        n = 48
        _check(kind, fast[:n], top(img[:n + r])[:n], f"{kind} top rows")
        _check(kind, fast[-n:], top(img[-(n + r):])[-n:], f"{kind} bottom rows")


def test_gigapixel_c5_gaussian_64bit_indexing():
    """BASELINE config c5 at full size: 32768 x 32768 RGB (3.2 GB > 2^31 bytes, which the reference's int
    arithmetic cannot address, image_filters.cu:95,:760), Gaussian sigma=5 radius=15.  Properties: the fused
    path equals the general path on the whole image; the oracle reproduces a band at the top edge, one that
    straddles the 2^31-byte offset, and one at the bottom edge."""
    import torch
    from gpu_image_processing_b200 import device
    L = _lib.load()
    H = W = 32768
    C, r, sigma = 3, 15, 5.0
    free, _ = torch.cuda.mem_get_info()
    if free < 14 * 2 ** 30:
        pytest.skip("needs ~13 GB of device memory")
    g = torch.Generator(device="cuda").manual_seed(42)
    x = torch.randint(0, 256, (H, W, C), dtype=torch.uint8, device="cuda", generator=g)
    fast = device.gaussian_blur(x, sigma, r, 2)
    old = L.gip_set_path(1)
    try:
        general = device.gaussian_blur(x, sigma, r, 2)
    finally:
        L.gip_set_path(old)
    assert torch.equal(fast, general)
    del general
    y_2g = (2 ** 31) // (W * C)                      # the row that contains byte offset 2^31
    for y0, y1 in ((0, 24), (y_2g - 12, y_2g + 12), (H - 24, H)):
        a, b = max(0, y0 - r), min(H, y1 + r)
        sub = x[a:b].cpu().numpy()
        want = O.gaussian_blur(sub, sigma, r)[y0 - a:y0 - a + (y1 - y0)]
        _check("gaussian", fast[y0:y1].cpu().numpy(), want, f"c5 rows {y0}:{y1}")


def test_frame_stream_c4_subset():
    """BASELINE config c4: 1920x1080 RGB frames, all three filters, one batched launch per filter; a 96-frame
    slice of the 4096-frame stream, every 16th frame against the oracle, fast path == general path on all."""
    import torch
    from gpu_image_processing_b200 import device
    L = _lib.load()
    g = torch.Generator(device="cuda").manual_seed(7)
    x = torch.randint(0, 256, (96, 1080, 1920, 3), dtype=torch.uint8, device="cuda", generator=g)
    runs = {"gaussian": lambda: device.gaussian_blur(x, 2.0, 3, 2), "box": lambda: device.box_blur(x, 3, 2),
            "sobel": lambda: device.sobel_edge_detection(x, 1)}
    for kind, run in runs.items():
        fast = run()
        old = L.gip_set_path(1)
        try:
            general = run()
        finally:
            L.gip_set_path(old)
        assert torch.equal(fast, general), kind
        for i in range(0, 96, 16):
            f = x[i].cpu().numpy()
            want = {"gaussian": lambda: O.gaussian_blur(f, 2.0, 3), "box": lambda: O.box_blur(f, 3),
                    "sobel": lambda: O.sobel(f, 1)}[kind]()
            _check(kind, fast[i].cpu().numpy(), want, f"c4 frame {i} {kind}")


@pytest.mark.parametrize("pin_in,pin_out", [(False, False), (True, False), (False, True), (True, True)])
def test_host_pipeline_chunked_large_buffers(pin_in, pin_out):
    """gip_*_host on buffers big enough for the chunked three-stream pipeline (and, for pageable caller memory,
    the staging threads): one 32 MB image in row chunks with halos, and a batch of frames in image chunks."""
    import torch
    L = _lib.load()
    m = _lib.Metrics()

    def buf(shape, pinned):
        t = torch.empty(shape, dtype=torch.uint8)
        return t.pin_memory() if pinned else t

    def run(kind, img, *args):
        x = buf(img.shape, pin_in)
        y = buf(img.shape, pin_out)
        x.numpy()[...] = img
        y.numpy()[...] = 0xAB
        if img.ndim == 4:
            b, h, w, c = img.shape
        else:
            (h, w, c), b = img.shape, 1
        fn = {"box": L.gip_box_blur_host, "gaussian": L.gip_gaussian_blur_host, "sobel": L.gip_sobel_host}[kind]
        _lib.check(fn(x.data_ptr(), y.data_ptr(), w, h, c, b, *args, ctypes.byref(m)))
        assert m.time_ms > 0
        return y.numpy().copy()

    img = synth.uniform(2048, 4096, 4, seed=99)
    _check("box", run("box", img, 5, 2), O.box_blur(img, 5), "box rows")
    _check("box", run("box", img, 31, 2), O.box_blur(img, 31), "box rows r31")
    _check("gaussian", run("gaussian", img, 2.0, 3, 1), O.gaussian_blur(img, 2.0, 3), "gaussian rows")
    _check("sobel", run("sobel", img, 1), O.sobel(img, 1), "sobel rows")
    one = synth.uniform(1080, 1920, 3, seed=17)                  # 6 MB: a single chunk, still through the staging threads
    _check("box", run("box", one, 7, 1), O.box_blur(one, 7), "box single chunk")
    _check("sobel", run("sobel", one, 2), O.sobel(one, 2), "sobel single chunk")
    frames = synth.uniform(24 * 540, 960, 3, seed=5).reshape(24, 540, 960, 3)
    got = run("box", frames, 3, 2)
    for i in (0, 7, 23):
        _check("box", got[i], O.box_blur(frames[i], 3), f"frame {i}")
    got = run("sobel", frames, 2)
    for i in (0, 11, 23):
        _check("sobel", got[i], O.sobel(frames[i], 2), f"frame {i}")


@pytest.mark.parametrize("c,w", [(3, 323), (1, 1001), (3, 640), (4, 257), (1, 64), (3, 2501), (4, 1027), (1, 5003), (3, 1000)])
@pytest.mark.parametrize("shift", [0, 1, 2, 3, 5])
def test_unaligned_buffers_and_canaries(c, w, shift, path):
    """Input and output at every byte alignment inside larger device buffers: the result still matches the oracle and
    not one byte outside the output image is written (the odd-pitch paths store aligned words that straddle
    neighbouring lanes' bytes).  The widest rows span several column strips of the fused kernels (1920 / 1968 bytes), so
    the strip seams of the any-alignment paths (overlapped consumer warps, chunks that straddle a row's ends) are covered."""
    import torch
    from gpu_image_processing_b200 import device
    h = 150
    img = synth.uniform(h, w, c, seed=w * 7 + c + shift)
    n = img.size
    pad = 64
    src = torch.zeros(n + 2 * pad, dtype=torch.uint8, device="cuda")
    x = src[pad + shift: pad + shift + n].view(h, w, c)
    x.copy_(torch.from_numpy(img))
    for kind, call, want in (
            ("box", lambda o: device.box_blur(x, 5, 2, out=o), O.box_blur(img, 5)),
            ("box", lambda o: device.box_blur(x, 20, 1, out=o), O.box_blur(img, 20)),
            ("gaussian", lambda o: device.gaussian_blur(x, 2.0, 3, 1, out=o), O.gaussian_blur(img, 2.0, 3)),
            ("gaussian", lambda o: device.gaussian_blur(x, 4.0, 9, 2, out=o), O.gaussian_blur(img, 4.0, 9)),
            ("sobel", lambda o: device.sobel_edge_detection(x, 1, out=o), O.sobel(img, 1)),
            ("sobel", lambda o: device.sobel_edge_detection(x, 2, out=o), O.sobel(img, 2))):
        dst = torch.full((n + 2 * pad,), 0xA5, dtype=torch.uint8, device="cuda")
        o = dst[pad + (shift * 3) % 7: pad + (shift * 3) % 7 + n].view(h, w, c)
        call(o)
        torch.cuda.synchronize()
        got = dst.cpu().numpy()
        lo = pad + (shift * 3) % 7
        assert (got[:lo] == 0xA5).all() and (got[lo + n:] == 0xA5).all(), (kind, "wrote outside the output image")
        _check(kind, got[lo:lo + n].reshape(h, w, c), want, f"{kind} shift={shift}")


@pytest.mark.parametrize("c", [1, 3, 4])
def test_in_place_calls(c, path):
    """d_output == d_input (and partially overlapping buffers): the reference's blurs tolerate it because they go
    through a temp image (image_filters.cu:760-880); here an overlapping input is saved to scratch first."""
    import torch
    from gpu_image_processing_b200 import device
    h, w = 300, 412
    img = synth.uniform(h, w, c, seed=c * 11)
    n = img.size
    for kind, call, want in (
            ("box", lambda x, o: device.box_blur(x, 9, 2, out=o), O.box_blur(img, 9)),
            ("gaussian", lambda x, o: device.gaussian_blur(x, 2.0, 3, 1, out=o), O.gaussian_blur(img, 2.0, 3)),
            ("sobel", lambda x, o: device.sobel_edge_detection(x, 1, out=o), O.sobel(img, 1))):
        x = torch.from_numpy(img).cuda()
        call(x, x)                                                   # exactly in place
        _check(kind, x.cpu().numpy(), want, kind + " in place")
        buf = torch.zeros(n + 4096, dtype=torch.uint8, device="cuda")   # output shifted 1024 bytes into the input
        xin = buf[:n].view(h, w, c)
        xin.copy_(torch.from_numpy(img))
        xout = buf[1024:1024 + n].view(h, w, c)
        call(xin, xout)
        _check(kind, xout.cpu().numpy(), want, kind + " overlapping")


@pytest.mark.parametrize("c", [1, 3, 4])
def test_bands_with_exactly_the_promised_halo_rows(c, path):
    """d_above / d_below hold rows_above / rows_below rows and not one more (gip_b200.h): one-row bands whose
    neighbours are separate, exactly sized device buffers.  (The Sobel load pipeline used to prefetch past them.)"""
    import torch
    L = _lib.load()
    h, w = 64, 1924
    img = synth.uniform(h, w, c, seed=5 + c)
    pitch = w * c
    cuts = [0, 30, 31, 32, 33, 64]
    stream = torch.cuda.current_stream().cuda_stream
    for kind, r in (("sobel", 1), ("box", 1), ("gaussian", 1)):
        bands = [torch.from_numpy(img[a:b].copy()).cuda() for a, b in zip(cuts[:-1], cuts[1:])]
        outs = [torch.zeros_like(t) for t in bands]
        for i, (y0, y1) in enumerate(zip(cuts[:-1], cuts[1:])):
            ra, rb = min(r, y0), min(r, h - y1)
            above = bands[i - 1].data_ptr() + (bands[i - 1].shape[0] - ra) * pitch if ra else None
            below = bands[i + 1].data_ptr() if rb else None
            args = (bands[i].data_ptr(), above, below, outs[i].data_ptr(), w, h, c, y0, y1 - y0, ra, rb)
            if kind == "sobel":
                rc = L.gip_sobel_band(*args, 2, stream)
            elif kind == "box":
                rc = L.gip_box_blur_band(*args, r, 1, stream)
            else:
                rc = L.gip_gaussian_blur_band(*args, 1.0, r, 1, stream)
            assert rc == 0
        torch.cuda.synchronize()
        got = np.concatenate([t.cpu().numpy() for t in outs], axis=0)
        want = {"sobel": lambda: O.sobel(img, 2), "box": lambda: O.box_blur(img, 1), "gaussian": lambda: O.gaussian_blur(img, 1.0, 1)}[kind]()
        _check(kind, got, want, kind + " one-row bands")


def test_c2_box_radius_sweep_as_benched():
    """BASELINE config c2 exactly as bench.py times it: box blur r = 1..31 on a 4096 x 4096 RGBA image, every radius
    against the oracle on row bands (top edge, an interior band, bottom edge)."""
    import torch
    from gpu_image_processing_b200 import device
    h = w = 4096
    c = 4
    img = synth.uniform(h, w, c, seed=2024)
    x = torch.from_numpy(img).cuda()
    y = torch.empty_like(x)
    n = 24
    for r in range(1, 32):
        device.box_blur(x, r, 2, out=y)
        out = y.cpu().numpy()
        _check("box", out[:n], O.box_blur(img[:n + r], r)[:n], f"c2 r={r} top rows")
        _check("box", out[-n:], O.box_blur(img[-(n + r):], r)[-n:], f"c2 r={r} bottom rows")
        _interior_band_check("box", img, out, h // 2 - 8 + r, h // 2 + 8 + r, r)


def test_first_call_time_ms_excludes_the_module_load():
    """The reference's time_ms is kernel time (image_filters.cu:804, :893-901): the first call that reaches a kernel
    variant must not report CUDA's lazy module load (tens of ms) as filter time.  Run in a fresh process."""
    import subprocess
    import sys
    code = ("import numpy as np, gpu_filters\n"
            "img = np.random.default_rng(1).integers(0, 256, (270, 480, 3), dtype=np.uint8)\n"
            "t = [gpu_filters.gaussian_blur(img, 2.0, 3, 1)['time_ms'], gpu_filters.box_blur(img, 3, 1)['time_ms'],\n"
            "     gpu_filters.sobel_edge_detection(img, 1)['time_ms']]\n"
            "print(max(t))\n")
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd=root, timeout=300)
    assert out.returncode == 0, out.stderr
    assert float(out.stdout.strip().splitlines()[-1]) < 2.0, out.stdout      # ms; a 389 KB image takes a few us


def test_two_devices_in_one_process():
    """include/image_filters.h: "the current device is whatever the caller set".  Shared-memory opt-ins, occupancy and
    SM counts are per-device state: every filter on cuda:0, then on cuda:1, in this one process (box r = 31 on RGBA
    needs > 48 KB of dynamic shared memory, Gaussian r = 9 takes the two-kernel path, r = 3 the fused one)."""
    import torch
    from gpu_image_processing_b200 import device
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two visible GPUs")
    img4 = synth.uniform(200, 1024, 4, seed=3)
    img3 = synth.uniform(200, 1024, 3, seed=4)
    want = [("box", O.box_blur(img4, 31)), ("box", O.box_blur(img3, 3)), ("gaussian", O.gaussian_blur(img3, 2.0, 3)),
            ("gaussian", O.gaussian_blur(img3, 3.0, 9)), ("gaussian", O.gaussian_blur(img3, 8.0, 24)),
            ("sobel", O.sobel(img3, 1)), ("sobel", O.sobel(img4, 2))]
    for dev in (0, 1, 0):
        with torch.cuda.device(dev):
            x4 = torch.from_numpy(img4).to(f"cuda:{dev}")
            x3 = torch.from_numpy(img3).to(f"cuda:{dev}")
            got = [device.box_blur(x4, 31, 2), device.box_blur(x3, 3, 1), device.gaussian_blur(x3, 2.0, 3, 1),
                   device.gaussian_blur(x3, 3.0, 9, 2), device.gaussian_blur(x3, 8.0, 24, 2),
                   device.sobel_edge_detection(x3, 1), device.sobel_edge_detection(x4, 2)]
            torch.cuda.synchronize(dev)
            for (kind, w_), g_ in zip(want, got):
                assert g_.device.index == dev
                _check(kind, g_.cpu().numpy(), w_, f"{kind} on cuda:{dev}")


@pytest.mark.parametrize("c", [1, 3, 4])
@pytest.mark.parametrize("r,sigma", [(16, 5.0), (19, 6.5), (24, 8.0), (27, 3.0), (31, 10.0)])
def test_gaussian_wide_radii_fast_path(c, r, sigma):
    """Radii 16..31 (the reference's level 2 accepts 2r+1 <= 64 taps, image_filters.cu:729) on the shift-formulation
    kernels: whole images of awkward sizes (narrower / shorter than the window, odd pitches), a batch, and row bands."""
    import torch
    from gpu_image_processing_b200 import device
    L = _lib.load()
    before = L.gip_launch_count()
    for h, w in ((1, 1), (3, 200), (200, 3), (45, 61), (130, 517), (300, 1025)):
        img = synth.uniform(h, w, c, seed=h * 7 + w + r)
        out = gpu_filters.gaussian_blur(img, sigma=sigma, radius=r, level=2)["image"]
        _check("gaussian", out, O.gaussian_blur(img, sigma, r), f"wide gaussian {h}x{w} c={c} r={r}")
    frames = np.stack([synth.smooth(150, 333, c, seed=s) for s in range(3)])
    got = device.gaussian_blur(torch.from_numpy(frames).cuda(), sigma, r, 1).cpu().numpy()
    for i in range(3):
        _check("gaussian", got[i], O.gaussian_blur(frames[i], sigma, r), f"wide gaussian batch c={c} r={r}")
    # row bands with halo pointers
    h, w = 260, 301
    img = synth.uniform(h, w, c, seed=r)
    x = torch.from_numpy(img).cuda()
    out = torch.zeros_like(x)
    pitch = w * c
    cuts = [0, 70, 131, 200, 260]
    stream = torch.cuda.current_stream().cuda_stream
    for y0, y1 in zip(cuts[:-1], cuts[1:]):
        ra, rb = min(r, y0), min(r, h - y1)
        rc = L.gip_gaussian_blur_band(x.data_ptr() + y0 * pitch, x.data_ptr() + (y0 - ra) * pitch if ra else None,
                                      x.data_ptr() + y1 * pitch if rb else None, out.data_ptr() + y0 * pitch, w, h, c,
                                      y0, y1 - y0, ra, rb, sigma, r, 1, stream)
        assert rc == 0
    torch.cuda.synchronize()
    _check("gaussian", out.cpu().numpy(), O.gaussian_blur(img, sigma, r), f"wide gaussian bands c={c} r={r}")
    # two launches (H, V) per call on the fast path; the general path would also be two, so check the kernel is the fast one
    assert L.gip_launch_count() - before >= 2 * (6 + 1 + 4)


@pytest.mark.parametrize("c,w", [(3, 239), (3, 241), (3, 243), (3, 481), (3, 721), (3, 7), (1, 239), (1, 241), (1, 483), (1, 1203), (1, 3)])
@pytest.mark.parametrize("shift", [0, 1, 3])
def test_sobel_any_alignment_strip_seams(c, w, shift):
    """Sobel on rows at any byte alignment, widths on either side of the 240-pixel strip seams and narrower than one
    lane: every strip -- the first and the last of a row included -- stages its row segments through the shared-memory
    ring and stores through the slab (csrc/fast_sobel.cu, VB = 1); both levels against the oracle, canaries around the
    output, a band height that is not a multiple of the 12-row tiles."""
    import torch
    from gpu_image_processing_b200 import device
    h = 37
    img = synth.uniform(h, w, c, seed=w * 3 + c + shift)
    n = img.size
    pad = 48
    src = torch.zeros(n + 2 * pad, dtype=torch.uint8, device="cuda")
    x = src[pad + shift: pad + shift + n].view(h, w, c)
    x.copy_(torch.from_numpy(img))
    for level in (1, 2):
        dst = torch.full((n + 2 * pad,), 0x5A, dtype=torch.uint8, device="cuda")
        lo = pad + (shift * 5) % 7
        o = dst[lo: lo + n].view(h, w, c)
        device.sobel_edge_detection(x, level, out=o)
        torch.cuda.synchronize()
        got = dst.cpu().numpy()
        assert (got[:lo] == 0x5A).all() and (got[lo + n:] == 0x5A).all(), "wrote outside the output image"
        _check("sobel", got[lo:lo + n].reshape(h, w, c), O.sobel(img, level), f"sobel seams level={level} shift={shift}")


def test_c1_shape_odd_pitch_sobel_and_box():
    """The reference's README shape (3239 x 2146 RGB, 9717-byte rows: no row but the first is 4-byte aligned) through Sobel
    and box blur: fast path == general path on the whole image, oracle on the top rows, an interior band, the bottom rows."""
    import torch
    from gpu_image_processing_b200 import device
    L = _lib.load()
    h, w, c = 2146, 3239, 3
    img = synth.uniform(h, w, c, seed=11)
    x = torch.from_numpy(img).cuda()
    for kind, run, orc, r in (("sobel", lambda: device.sobel_edge_detection(x, 1), lambda s: O.sobel(s, 1), 1),
                              ("sobel", lambda: device.sobel_edge_detection(x, 2), lambda s: O.sobel(s, 2), 1),
                              ("box", lambda: device.box_blur(x, 3, 2), lambda s: O.box_blur(s, 3), 3),
                              ("box", lambda: device.box_blur(x, 12, 1), lambda s: O.box_blur(s, 12), 12)):
        fast = run().cpu().numpy()
        old = L.gip_set_path(1)
        try:
            general = run().cpu().numpy()
        finally:
            L.gip_set_path(old)
        assert np.array_equal(fast, general), kind
        n = 40
        _check(kind, fast[:n], orc(img[:n + r])[:n], f"c1-shape {kind} top rows")
        _check(kind, fast[-n:], orc(img[-(n + r):])[-n:], f"c1-shape {kind} bottom rows")
        y0 = h // 2
        _check(kind, fast[y0:y0 + n], orc(img[y0 - r:y0 + n + r])[r:r + n], f"c1-shape {kind} interior rows")
