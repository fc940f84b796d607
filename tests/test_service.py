"""In-process tests of the REST contract (reference: backend/app.py) and of the ncu side-car's parser."""
import base64
import io

import numpy as np
import pytest
from fastapi.testclient import TestClient
from PIL import Image

from gpu_image_processing_b200.profiling import ncu_profiler
from gpu_image_processing_b200.service.app import app
from tests import synth

client = TestClient(app)


def _png_b64(arr, prefix=True):
    buf = io.BytesIO()
    Image.fromarray(arr).save(buf, format="PNG")
    s = base64.b64encode(buf.getvalue()).decode()
    return ("data:image/png;base64," + s) if prefix else s


def _decode(data_url):
    return np.array(Image.open(io.BytesIO(base64.b64decode(data_url.split(",")[1]))))


def test_static_routes():
    assert client.get("/").json()["name"] == "GPU Image Processing API"
    assert client.get("/api/health").json()["status"] == "healthy"
    f = client.get("/api/filters").json()["filters"]
    assert set(f) == {"gaussian", "box", "sobel"} and f["gaussian"]["parameters"]["sigma"]["default"] == 2.0


def test_bad_filter_and_level_are_400():
    img = _png_b64(synth.uniform(8, 8, 3))
    r = client.post("/api/process", json={"image": img, "filter": "median"})
    assert r.status_code == 400 and "Invalid filter" in r.json()["detail"]
    r = client.post("/api/process", json={"image": img, "filter": "box", "level": 3})
    assert r.status_code == 400 and "Invalid level: 3" in r.json()["detail"]


def test_undecodable_image_is_wrapped_as_500():
    r = client.post("/api/process", json={"image": "not-base64!!", "filter": "sobel", "level": 1})
    assert r.status_code == 500 and "Processing failed" in r.json()["detail"]      # app.py:283-284


@pytest.mark.gpu
def test_process_round_trip_matches_oracle():
    from oracle import oracle as O
    rgb = synth.smooth(97, 131, 3, seed=3)
    gray = synth.uniform(40, 50, 1)[:, :, 0]
    for payload, want in (
            ({"filter": "gaussian", "level": 2, "sigma": 1.5, "radius": 4}, O.gaussian_blur(rgb, 1.5, 4)),
            ({"filter": "box", "level": 1, "radius": 5}, O.box_blur(rgb, 5)),
            ({"filter": "sobel", "level": 1}, O.sobel(rgb, 1)),
            ({"filter": "sobel", "level": 2}, O.sobel(rgb, 2))):
        r = client.post("/api/process", json=dict(payload, image=_png_b64(rgb, prefix=payload["level"] == 1)))
        assert r.status_code == 200, r.text
        body = r.json()
        assert np.array_equal(_decode(body["processed_image"]), want)
        assert set(body["metrics"]) == {"time_ms", "bandwidth_gbps", "fps"} and body["metrics"]["time_ms"] > 0
        assert body["info"]["width"] == 131 and body["info"]["height"] == 97 and body["info"]["channels"] == 3
    # grayscale uploads are converted to RGB first (app.py:80-83)
    r = client.post("/api/process", json={"image": _png_b64(gray), "filter": "box", "radius": 2})
    assert r.status_code == 200 and r.json()["info"]["channels"] == 3
    both = client.post("/api/process-all", json={"image": _png_b64(rgb), "filter": "sobel"}).json()
    assert set(both["results"]) == {"level_1", "level_2"} and both["results"]["level_2"]["info"]["level_number"] == 2
    assert both["results"]["level_1"]["info"]["level"] == "naive" and both["profiling_available"] is False


RAW = '''"ID","Process ID","Process Name","Host Name","Kernel Name","Context","Stream","Block Size","Grid Size","Device","CC","gpu__time_duration.sum","dram__bytes_read.sum","dram__bytes_write.sum","sm__warps_active.avg.pct_of_peak_sustained_active","launch__block_size","launch__grid_size","sm__throughput.avg.pct_of_peak_sustained_elapsed"
"","","","","","","","","","","","usecond","Mbyte","Mbyte","%","","","%"
"0","1","python","h","gip_gauss_h<3, 3, 1>(Job)","1","7","(192, 1, 1)","(10, 1, 1)","0","10.0","120.5","99.5","50.0","17.8","192","10","40.0"
"1","1","python","h","gip_gauss_v<3>(Job)","1","7","(128, 1, 1)","(20, 1, 1)","0","10.0","100.5","100.0","50.0","49.5","128","20","35.0"
'''


def test_ncu_csv_parser_keeps_the_reference_shape():
    m = ncu_profiler.parse_ncu_raw_csv(RAW)
    assert set(m) >= {"occupancy", "memory", "warp", "execution", "throughput", "config", "kernel_durations",
                      "total_kernel_duration_ms", "kernels_profiled", "total_kernels"}
    assert m["total_kernels"] == 2 and m["total_kernel_duration_ms"] == pytest.approx(0.221)
    common = ncu_profiler.get_common_ncu_metrics(m, ncu_data=m)
    assert common["time_ms"] == pytest.approx(0.221) and common["total_kernels"] == 2
    assert common["occupancy_pct"] == pytest.approx(49.5)
    assert ncu_profiler.get_common_ncu_metrics({}) == {}
    with pytest.raises(ValueError):
        ncu_profiler.profile_kernel_with_ncu(np.zeros((4, 4, 3), np.uint8), "median", 1)


def test_upload_route_returns_base64_and_shape():
    """POST /api/upload (reference backend/app.py:496-524)."""
    arr = synth.uniform(9, 13, 3, seed=3)
    buf = io.BytesIO()
    Image.fromarray(arr).save(buf, format="PNG")
    r = client.post("/api/upload", files={"file": ("x.png", buf.getvalue(), "image/png")})
    assert r.status_code == 200
    j = r.json()
    assert (j["width"], j["height"], j["channels"]) == (13, 9, 3)
    assert np.array_equal(_decode(j["base64_image"]), arr)
    gray = Image.fromarray(arr[:, :, 0])                 # mode "L" stays gray; "channels" is the reference's quirk (ndim)
    buf = io.BytesIO(); gray.save(buf, format="PNG")
    j = client.post("/api/upload", files={"file": ("g.png", buf.getvalue(), "image/png")}).json()
    assert j["channels"] == 2
    r = client.post("/api/upload", files={"file": ("x.png", b"not an image", "image/png")})
    assert r.status_code == 500 and "Upload failed" in r.json()["detail"]


def test_process_all_undecodable_image_is_500_like_the_reference():
    r = client.post("/api/process-all", json={"image": "not-base64!!", "filter": "sobel"})
    assert r.status_code == 500 and "Processing failed" in r.json()["detail"]        # app.py:493-494


def test_ncu_parser_sums_band_launches_per_call():
    """The host path cuts big images into row-band chunks (one launch each): durations and DRAM bytes are summed over the
    launches of a call, not averaged per kernel name."""
    rows = RAW.strip().split("\n")
    text = "\n".join(rows[:2] + [rows[2]] * 6 + [rows[3]] * 6) + "\n"      # 3 calls x 2 chunks x (H, V)
    m = ncu_profiler.parse_ncu_raw_csv(text, calls=3)
    assert m["total_kernel_duration_ms"] == pytest.approx(2 * 0.221)
    assert m["launches_per_call"] == {"gip_gauss_h<3, 3, 1>": 2.0, "gip_gauss_v<3>": 2.0}
    assert m["memory"]["dram_bytes_per_call"] == pytest.approx(2 * (99.5 + 50 + 100 + 50) * 1e6)
