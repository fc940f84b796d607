"""The drop-in claim under test: the reference's OWN pybind11 binding (backend/cuda_bindings/bindings.cpp, unmodified,
compiled from where it lies under /root/reference) builds against this repo's include/image_filters.h, links against
libgip_b200.so through the three mangled C++ entry points, and imports as `gpu_filters`.

CPU test: compile + import (skipped where /root/reference is not mounted).
GPU test: the prebuilt module from `make -C oracle dropin` (oracle/_ref/dropin/, travels to the GPU box) runs the three
filters through the reference's binding code and the results equal the oracle's.  No reference source is copied."""
import os
import subprocess
import sys
import sysconfig

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_BINDING = "/root/reference/backend/cuda_bindings/bindings.cpp"
LIBDIR = os.path.join(ROOT, "gpu_image_processing_b200")
DROPIN_DIR = os.path.join(ROOT, "oracle", "_ref", "dropin")


def _run(code, path0):
    env = dict(os.environ)
    env["PYTHONPATH"] = path0 + os.pathsep + ROOT
    return subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd=path0, env=env, timeout=600)


@pytest.mark.skipif(not os.path.exists(REF_BINDING), reason="reference sources are not mounted")
def test_reference_binding_compiles_links_and_imports(tmp_path):
    import pybind11
    out = tmp_path / ("gpu_filters" + sysconfig.get_config_var("EXT_SUFFIX"))
    cmd = ["g++", "-O0", "-shared", "-fPIC", "-std=c++17", "-I" + os.path.join(ROOT, "include"), "-I/usr/local/cuda/include",
           "-I" + pybind11.get_include(), "-I" + sysconfig.get_paths()["include"], REF_BINDING, "-o", str(out),
           "-L" + LIBDIR, "-l:libgip_b200.so", "-L/usr/local/cuda/lib64", "-lcudart_static", "-ldl", "-lrt", "-lpthread",
           "-Wl,-rpath," + LIBDIR]
    res = subprocess.run(cmd, capture_output=True, text=True)
    assert res.returncode == 0, res.stderr[-2000:]
    code = ("import gpu_filters as g, inspect\n"
            "assert g.__file__.endswith('.so'), g.__file__\n"
            "assert (g.NAIVE, g.SHARED_MEMORY, g.TEXTURE_MEMORY) == (1, 2, 3)\n"
            "for f in ('gaussian_blur', 'box_blur', 'sobel_edge_detection'): assert callable(getattr(g, f))\n"
            "print('ok')\n")
    res = _run(code, str(tmp_path))
    assert res.returncode == 0 and "ok" in res.stdout, res.stderr[-2000:]
    # the only undefined non-libc/cudart symbols of the binding are the three entry points this library exports
    nm = subprocess.run(["nm", "-D", "--undefined-only", str(out)], capture_output=True, text=True).stdout
    wanted = [l.split()[-1] for l in nm.splitlines() if "OptimizationLevel" in l]
    assert len(wanted) == 3, wanted
    have = subprocess.run(["nm", "-D", "--defined-only", os.path.join(LIBDIR, "libgip_b200.so")], capture_output=True, text=True).stdout
    for sym in wanted:
        assert sym in have, sym


@pytest.mark.gpu
def test_reference_binding_runs_on_this_library_and_matches_the_oracle():
    if not (os.path.isdir(DROPIN_DIR) and any(f.startswith("gpu_filters") and f.endswith(".so") for f in os.listdir(DROPIN_DIR))):
        pytest.skip("oracle/_ref/dropin was not built (make -C oracle dropin, needs /root/reference)")
    code = (
        "import numpy as np, gpu_filters as g\n"
        "assert g.__file__.endswith('.so') and 'dropin' in g.__file__, g.__file__\n"
        "from oracle import oracle as O\n"
        "from tests import synth\n"
        "for c in (1, 3, 4):\n"
        "    img = synth.uniform(211, 333, c, seed=c)\n"
        "    for lvl in (1, 2):\n"
        "        r = g.gaussian_blur(img, sigma=2.0, radius=3, level=lvl)\n"
        "        assert set(r) == {'image', 'time_ms', 'bandwidth_gbps', 'fps'} and r['time_ms'] > 0\n"
        "        assert np.array_equal(r['image'], O.gaussian_blur(img, 2.0, 3)), ('gaussian', c, lvl)\n"
        "        assert np.array_equal(g.box_blur(img, radius=7, level=lvl)['image'], O.box_blur(img, 7)), ('box', c, lvl)\n"
        "        assert np.array_equal(g.sobel_edge_detection(img, level=lvl)['image'], O.sobel(img, lvl)), ('sobel', c, lvl)\n"
        "try:\n"
        "    g.box_blur(img, radius=3, level=3)\n"
        "    raise SystemExit('level 3 accepted')\n"
        "except RuntimeError:\n"
        "    pass\n"
        "print('ok')\n")
    res = _run(code, DROPIN_DIR)
    assert res.returncode == 0 and "ok" in res.stdout, (res.stdout[-1000:], res.stderr[-3000:])
