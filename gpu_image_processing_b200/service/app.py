"""FastAPI re-host of the reference's REST contract (backend/app.py) on the B200 library.

Same routes, request/response models, status codes and texts:
    GET  /                 app.py:115-130      GET /api/health  app.py:132-138     GET /api/filters  app.py:140-184
    POST /api/process      app.py:186-284      body {image, filter, level=1, sigma=2.0, radius=3, enable_profiling=false}
                                               -> {processed_image: "data:image/png;base64,...", metrics{time_ms,bandwidth_gbps,fps}, info{...}}
                                               503 module missing, 400 bad filter / level, 500 everything else
    POST /api/process-all  app.py:286-494      both levels + optional ncu metrics, keys level_1 / level_2; a level that fails is
                                               logged and skipped, 500 only when no level succeeded (app.py:461-475)
    POST /api/upload       app.py:496-524      multipart file -> {base64_image, width, height, channels}
Images are decoded with PIL and forced to RGB like the reference (app.py:80-83).  Handlers are plain `def`:
FastAPI runs them in its thread pool, so a long GPU call no longer blocks the event loop (the reference's
handlers are `async def` but fully blocking, app.py:187).
"""
from __future__ import annotations

import base64
import contextlib
import io
from typing import Any, Dict, Optional

import numpy as np
from fastapi import FastAPI, File, HTTPException, UploadFile
from fastapi.middleware.cors import CORSMiddleware
from PIL import Image
from pydantic import BaseModel

try:
    from .. import gpu_filters
    from .._lib import load as _load_lib
    _load_lib()
    GPU_AVAILABLE = True
except Exception as e:  # library not built: same degraded mode as the reference (app.py:21-27)
    print(f"Warning: GPU filters module not found: {e}")
    gpu_filters = None
    GPU_AVAILABLE = False

def _warm_up():
    """First use of a kernel pays CUDA's lazy module load (tens of ms, and it would show up in the first request's
    time_ms).  Run each filter once on a tiny image at start-up; a box without a GPU just skips it."""
    if gpu_filters is None:
        return
    try:
        tiny = np.zeros((32, 32, 3), np.uint8)
        gpu_filters.gaussian_blur(tiny)
        gpu_filters.box_blur(tiny)
        gpu_filters.sobel_edge_detection(tiny)
    except Exception:       # no device, wrong driver: the request handlers report it
        pass


@contextlib.asynccontextmanager
async def _lifespan(_app):
    _warm_up()
    yield


app = FastAPI(title="GPU Image Processing API", description="High-performance CUDA-accelerated image processing filters",
              version="1.0.0", lifespan=_lifespan)
app.add_middleware(CORSMiddleware, allow_origins=["*"], allow_credentials=True, allow_methods=["*"], allow_headers=["*"])


class FilterRequest(BaseModel):
    image: str
    filter: str
    level: int = 1
    sigma: Optional[float] = 2.0
    radius: Optional[int] = 3
    enable_profiling: bool = False


class FilterResponse(BaseModel):
    processed_image: str
    metrics: Dict[str, Any]
    info: Dict[str, Any]


class AllLevelsResponse(BaseModel):
    original_image: str
    results: Dict[str, FilterResponse]
    image_info: Dict[str, Any]
    profiling_available: bool = False


def decode_base64_image(base64_str: str) -> np.ndarray:
    try:
        if "," in base64_str:
            base64_str = base64_str.split(",")[1]
        image = Image.open(io.BytesIO(base64.b64decode(base64_str)))
        if image.mode != "RGB":
            image = image.convert("RGB")
        return np.array(image)
    except Exception as e:
        raise HTTPException(status_code=400, detail=f"Failed to decode image: {str(e)}")


def encode_image_to_base64(img_array: np.ndarray) -> str:
    try:
        buf = io.BytesIO()
        Image.fromarray(np.ascontiguousarray(img_array, dtype=np.uint8)).save(buf, format="PNG", compress_level=1)
        return "data:image/png;base64," + base64.b64encode(buf.getvalue()).decode("utf-8")
    except Exception as e:
        raise HTTPException(status_code=500, detail=f"Failed to encode image: {str(e)}")


LEVEL_NAMES = {"gaussian": {1: "naive", 2: "texture_memory"}, "box": {1: "naive", 2: "shared_memory"},
               "sobel": {1: "naive", 2: "shared_memory"}}
LEVEL_ERRORS = {
    "gaussian": "Gaussian blur supports levels 1 (naive) or 2 (texture_memory)",
    "box": "Box blur supports levels 1 (naive) or 2 (shared_memory)",
    "sobel": "Sobel edge detection supports levels 1 (naive) or 2 (shared_memory)",
}


def _apply(req: FilterRequest, img: np.ndarray, level: int):
    if req.filter == "gaussian":
        return gpu_filters.gaussian_blur(img, sigma=req.sigma, radius=req.radius, level=level)
    if req.filter == "box":
        return gpu_filters.box_blur(img, radius=req.radius, level=level)
    return gpu_filters.sobel_edge_detection(img, level=level)


def _info(req: FilterRequest, level: int, shape, with_number: bool = False) -> Dict[str, Any]:
    h, w, c = shape
    info = {"filter": req.filter, "level": LEVEL_NAMES[req.filter][level], "width": int(w), "height": int(h),
            "channels": int(c),
            "parameters": {"sigma": req.sigma if req.filter == "gaussian" else None,
                           "radius": req.radius if req.filter in ("gaussian", "box") else None}}
    if with_number:
        info["level_number"] = level
    return info


def _check(req: FilterRequest):
    if not GPU_AVAILABLE:
        raise HTTPException(status_code=503, detail="GPU filters module not available. Build it first.")
    if req.filter not in ("gaussian", "box", "sobel"):
        raise HTTPException(status_code=400, detail=f"Invalid filter: {req.filter}. Must be 'gaussian', 'box', or 'sobel'")


@app.get("/")
def root():
    return {"name": "GPU Image Processing API", "version": "1.0.0", "status": "running", "gpu_available": GPU_AVAILABLE,
            "endpoints": {"GET /": "This message", "GET /api/filters": "List available filters",
                          "POST /api/process": "Process image with filter", "GET /api/health": "Health check"}}


@app.get("/api/health")
def health_check():
    return {"status": "healthy", "gpu_available": GPU_AVAILABLE}


@app.get("/api/filters")
def list_filters():
    lvl = {"type": "int", "default": 1, "options": [1, 2]}
    return {"filters": {
        "gaussian": {"name": "Gaussian Blur", "description": "Smooth blur with weighted averaging (bell curve)",
                     "parameters": {"sigma": {"type": "float", "default": 2.0, "range": [0.5, 20.0]},
                                    "radius": {"type": "int", "default": 3, "range": [1, 15]}, "level": lvl},
                     "optimization_levels": {"1": "Naive (global memory)", "2": "Texture + Constant + Vectorized (optimized)"}},
        "box": {"name": "Box Blur", "description": "Simple average blur (faster than Gaussian)",
                "parameters": {"radius": {"type": "int", "default": 3, "range": [1, 15]}, "level": lvl},
                "optimization_levels": {"1": "Naive (global memory)", "2": "Shared memory tiling"}},
        "sobel": {"name": "Sobel Edge Detection", "description": "Detect edges using gradient magnitude (Gx, Gy)",
                  "parameters": {"level": {"type": "int", "default": 2, "options": [1, 2]}},
                  "optimization_levels": {"1": "Naive (global memory)", "2": "Shared memory (18x+ faster)"}}},
        "gpu_available": GPU_AVAILABLE}


@app.post("/api/process", response_model=FilterResponse)
def process_image(request: FilterRequest):
    _check(request)
    if request.level not in (1, 2):
        raise HTTPException(status_code=400, detail=f"Invalid level: {request.level}. {LEVEL_ERRORS[request.filter]}")
    try:
        img = decode_base64_image(request.image)
        result = _apply(request, img, request.level)
        return FilterResponse(processed_image=encode_image_to_base64(result["image"]),
                              metrics={"time_ms": float(result["time_ms"]), "bandwidth_gbps": float(result["bandwidth_gbps"]),
                                       "fps": float(result["fps"])},
                              info=_info(request, request.level, img.shape))
    except Exception as e:      # includes the 400 from decode, exactly like the reference (app.py:283-284)
        raise HTTPException(status_code=500, detail=f"Processing failed: {str(e)}")


@app.post("/api/process-all", response_model=AllLevelsResponse)
def process_all_levels(request: FilterRequest):
    _check(request)
    try:
        img = decode_base64_image(request.image)
        profiling = False
        if request.enable_profiling:
            from ..profiling import check_ncu_available
            profiling = check_ncu_available()
        results: Dict[str, FilterResponse] = {}
        for level in (1, 2):
            try:
                result = _apply(request, img.copy(), level)
                metrics: Dict[str, Any] = {"time_ms": float(result["time_ms"]), "bandwidth_gbps": float(result["bandwidth_gbps"]),
                                           "fps": float(result["fps"])}
                if profiling:
                    try:
                        from ..profiling import get_common_ncu_metrics, profile_kernel_with_ncu
                        ncu = profile_kernel_with_ncu(img.copy(), request.filter, level,
                                                      request.sigma if request.filter == "gaussian" else None,
                                                      request.radius if request.filter in ("gaussian", "box") else None)
                        common = get_common_ncu_metrics(ncu, ncu_data=ncu)
                        if common.get("time_ms", 0) > 0:
                            metrics["ncu_profiled_time_ms"] = common["time_ms"]
                        metrics.update({k: v for k, v in common.items() if k != "time_ms"})   # CUDA-event time stays primary
                        metrics["ncu_data"] = ncu
                    except Exception as e:
                        metrics["profiling_error"] = str(e)
                results[f"level_{level}"] = FilterResponse(processed_image=encode_image_to_base64(result["image"]),
                                                           metrics=metrics, info=_info(request, level, img.shape, True))
            except Exception as e:      # app.py:461-466: log, keep going with the other level
                print(f"Error processing level {level}: {str(e)}")
                continue
        if not results:                 # app.py:468-475
            raise HTTPException(status_code=500, detail="Failed to process image with any optimization level")
        h, w, c = img.shape
        return AllLevelsResponse(original_image=encode_image_to_base64(img), results=results,
                                 image_info={"width": int(w), "height": int(h), "channels": int(c), "filter": request.filter,
                                             "parameters": {"sigma": request.sigma if request.filter == "gaussian" else None,
                                                            "radius": request.radius if request.filter in ("gaussian", "box") else None}},
                                 profiling_available=profiling)
    except Exception as e:              # every failure, HTTP errors included, leaves as 500 like the reference (app.py:493-494)
        raise HTTPException(status_code=500, detail=f"Processing failed: {str(e)}")


@app.post("/api/upload")
def upload_image(file: UploadFile = File(...)):
    """Upload an image file, get it back base64-encoded (app.py:496-524)."""
    try:
        image = Image.open(io.BytesIO(file.file.read()))
        if image.mode not in ("RGB", "L"):
            image = image.convert("RGB")
        arr = np.array(image)
        return {"base64_image": encode_image_to_base64(arr), "width": image.width, "height": image.height,
                "channels": len(arr.shape) if len(arr.shape) == 2 else arr.shape[2]}
    except Exception as e:
        raise HTTPException(status_code=500, detail=f"Upload failed: {str(e)}")


if __name__ == "__main__":
    import uvicorn
    uvicorn.run(app, host="0.0.0.0", port=8000)
