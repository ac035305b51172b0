"""ctypes binding of libof_b200.so (include/of_b200.h) -- the only way the drop-in modules
reach the GPU.  There is no CPU fallback: if the library is not built or no CUDA device is
present, every compute call raises.

Host side of the drop-in boundary: NumPy arrays in, NumPy arrays out, same argument
meaning as the reference's functions (python/lucas_kanade_core.py,
python/lucas_kanade_pyramidal.py).  ``*_dev`` helpers take raw device pointers (e.g.
``torch.Tensor.data_ptr()``) for batched, device-resident use.
"""

from __future__ import annotations

import ctypes as C
import os
from pathlib import Path
from typing import Optional, Sequence, Tuple

import numpy as np

HERE = Path(__file__).resolve().parent
LIB_PATH = HERE / os.environ.get("OF_B200_LIB_NAME", "libof_b200.so")  # variants: experiments only

MODE_EXACT = 0
MODE_FAST = 1
FX_MIRROR_AVG_QUIRK = 1

_STATUS = {
    1: "invalid argument",
    2: "CUDA error",
    3: "unsupported",
    4: "no CUDA device",
    5: "out of device memory",
    6: "peer time-out",
}


class OFBackendError(RuntimeError):
    """The CUDA backend failed (or is missing).  Never silently replaced by CPU code."""


_lib = None

_f32p = C.POINTER(C.c_float)
_f64p = C.POINTER(C.c_double)
_u8p = C.POINTER(C.c_uint8)
_i16p = C.POINTER(C.c_int16)
_i32p = C.POINTER(C.c_int)
_vp = C.c_void_p
_i = C.c_int

# name -> (restype, argtypes); also the list the ABI test checks against include/of_b200.h
SIGNATURES = {
    "of_version": (_i, []),
    "of_last_error": (C.c_char_p, []),
    "of_device_count": (_i, []),
    "of_set_device": (_i, [_i]),
    "of_kernel_launches": (C.c_longlong, []),
    "of_host_alloc_pinned": (_i, [C.POINTER(_vp), C.c_size_t]),
    "of_host_free_pinned": (_i, [_vp]),
    "of_release_host_buffers": (_i, []),
    "of_gradients_f32": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i]),
    "of_lk_from_gradients_f32": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i]),
    "of_lk_single_scale_f32": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i]),
    "of_pyramid_down_f32": (_i, [_vp, _vp, _i, _i, _i, _i, _vp, _i]),
    "of_warp_f32": (_i, [_vp, _vp, _vp, _vp, _i, _i]),
    "of_upsample_flow_f32": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i]),
    "of_lk_pyramidal_f32": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _vp, _i, _vp, _vp]),
    "of_lk_single_scale_fx": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i]),
    "of_lk_single_scale_f32_dev": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp]),
    "of_lk_pyramidal_workspace_bytes": (C.c_size_t, [_i, _i, _i, _i, _i]),
    "of_lk_pyramidal_f32_dev": (
        _i,
        [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _vp, _i, _vp, C.c_size_t, _vp, _vp, _vp],
    ),
    "of_lk_single_scale_fx_dev": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp]),
    "of_pyramid_down_f32_dev": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _vp, _i, _i, _i, _i, _vp]),
    "of_lk_refine_pingpong_f32_dev": (
        _i,
        [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _i, _vp, _vp, C.c_size_t, _vp],
    ),
    "of_lk_convergence_update_dev": (_i, [_vp, _i, C.c_double, _vp, _vp, _vp, _vp, _i, _i, _vp]),
    "of_upsample_flow_f32_dev": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _vp]),
    "of_lk_refine_workspace_bytes": (C.c_size_t, [_i, _i, _i]),
    "of_lk_refine_f32_dev": (
        _i,
        [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _i, _vp, _vp, C.c_size_t, _vp],
    ),
    "of_lk_single_scale_u8": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i]),
    "of_lk_single_scale_u8_dev": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp]),
    "of_load_frame_u8": (_i, [C.c_char_p, _vp, _i, _i]),
    "of_export_flow_txt": (_i, [C.c_char_p, _vp, _vp, _i, _i, _i, _i, _i, _i]),
    "of_export_flow_fx_txt": (_i, [C.c_char_p, _vp, _vp, _i, _i, _i, _i, _i, _i]),
    "of_apply_motion_u8": (_i, [_vp, _vp, _i, _i, _i, _vp, _vp, C.c_double]),
    "of_apply_motion_u8_dev": (_i, [_vp, _vp, _i, _i, _i, _vp, _vp, C.c_double, _vp]),
    "of_warp_affine_u8": (_i, [_vp, _vp, _i, _i, _i, _vp, _i]),
    "of_warp_affine_u8_dev": (_i, [_vp, _vp, _i, _i, _i, _vp, _i, _vp]),
    "of_flow_metrics_workspace_bytes": (C.c_size_t, [_i, _i, _i]),
    "of_flow_metrics_f32_dev": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _vp, _vp, C.c_size_t, _vp]),
    "of_flow_metrics_f32": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _vp]),
    "of_rowband_create": (_i, [C.POINTER(_vp), _i, _i, _i, _i, _i, _i, _i, _i, _vp, _i]),
    "of_rowband_arena_bytes": (C.c_size_t, [_vp]),
    "of_rowband_arena": (_vp, [_vp]),
    "of_rowband_ipc_handle": (_i, [_vp, _vp]),
    "of_rowband_open_peers_ipc": (_i, [_vp, _vp]),
    "of_rowband_set_peers": (_i, [_vp, C.POINTER(_vp)]),
    "of_rowband_set_replicate_pixels": (_i, [_vp, C.c_longlong]),
    "of_rowband_set_timeout_ms": (_i, [_vp, _i]),
    "of_rowband_run": (_i, [_vp, _vp, _vp, _vp, _vp, _vp]),
    "of_rowband_result": (_i, [_vp, C.POINTER(_vp), C.POINTER(_vp)]),
    "of_rowband_trace": (_i, [_vp, _vp, _vp, _vp, _vp]),
    "of_rowband_status": (_i, [_vp, _vp]),
    "of_rowband_destroy": (_i, [_vp]),
}


def lib() -> C.CDLL:
    """Load libof_b200.so once.  Raises OFBackendError when it has not been built."""
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise OFBackendError(
                f"{LIB_PATH} is missing: build it with `python optical-flow-fpga_b200/build.py` "
                "(nvcc, sm_100a).  This backend has no CPU fallback."
            )
        handle = C.CDLL(str(LIB_PATH))
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def _check(status: int) -> None:
    if status != 0:
        msg = lib().of_last_error().decode("utf-8", "replace")
        kind = _STATUS.get(status, f"status {status}")
        if status == 1:
            raise ValueError(f"of_b200: {msg}")
        raise OFBackendError(f"of_b200 {kind}: {msg}")


def device_count() -> int:
    return int(lib().of_device_count())


def set_device(ordinal: int) -> None:
    _check(lib().of_set_device(int(ordinal)))


def kernel_launches() -> int:
    return int(lib().of_kernel_launches())


def default_mode() -> int:
    """EXACT unless OF_B200_MODE=fast: parity with the reference comes first."""
    return MODE_FAST if os.environ.get("OF_B200_MODE", "exact").lower() == "fast" else MODE_EXACT


def _frame(a, name: str) -> np.ndarray:
    arr = np.ascontiguousarray(a, dtype=np.float32)
    if arr.ndim != 2:
        raise ValueError(f"{name} must be a 2-D array, got shape {arr.shape}")
    return arr


def _out_pair(out, shape, dtype=np.float32):
    """Caller-supplied result buffers of a batched call: the C library writes through their raw pointers,
    so anything but two distinct C-contiguous arrays of exactly `shape` / `dtype` would be an out-of-bounds
    host write."""
    try:
        u, v = out
    except Exception:
        raise ValueError("out must be a pair (u, v) of arrays") from None
    for name, a in (("out[0]", u), ("out[1]", v)):
        if not isinstance(a, np.ndarray) or a.dtype != np.dtype(dtype) or tuple(a.shape) != tuple(shape) \
                or not a.flags["C_CONTIGUOUS"] or not a.flags["WRITEABLE"]:
            raise ValueError(f"{name} must be a writable C-contiguous {np.dtype(dtype).name} array of shape {tuple(shape)}")
    if u is v or np.shares_memory(u, v):
        raise ValueError("out[0] and out[1] must not overlap")
    return u, v


def release_host_buffers() -> None:
    """Free the device buffers the host-buffer calls keep for the current device (re-allocated on demand)."""
    _check(lib().of_release_host_buffers())


def _same_shape(*arrs: np.ndarray) -> None:
    s = arrs[0].shape
    for a in arrs[1:]:
        if a.shape != s:
            raise ValueError(f"shape mismatch: {s} vs {a.shape}")


def _ptr(a: np.ndarray) -> int:
    return a.ctypes.data


def _window(window_size: int) -> int:
    w = int(window_size)
    if w < 1 or w % 2 == 0:
        raise ValueError("window_size must be odd")
    return w


def gaussian_weights(sigma: float, truncate: float = 4.0) -> np.ndarray:
    """The float64 taps scipy.ndimage.gaussian_filter uses for this sigma, built with the
    same NumPy expression so the kernel multiplies by bit-identical weights."""
    radius = int(truncate * float(sigma) + 0.5)
    x = np.arange(-radius, radius + 1)
    phi = np.exp(-0.5 / (float(sigma) * float(sigma)) * x**2)
    return np.ascontiguousarray(phi / phi.sum(), dtype=np.float64)


# --------------------------------------------------------------------------------------
# host-array API (NumPy in / NumPy out)
# --------------------------------------------------------------------------------------
def gradients(frame_prev, frame_curr):
    p, c = _frame(frame_prev, "frame_prev"), _frame(frame_curr, "frame_curr")
    _same_shape(p, c)
    h, w = p.shape
    ix, iy, it = (np.empty((h, w), np.float32) for _ in range(3))
    _check(lib().of_gradients_f32(_ptr(p), _ptr(c), _ptr(ix), _ptr(iy), _ptr(it), h, w))
    return ix, iy, it


def lk_from_gradients(ix, iy, it, window_size: int = 5):
    gx, gy, gt = _frame(ix, "Ix"), _frame(iy, "Iy"), _frame(it, "It")
    _same_shape(gx, gy, gt)
    h, w = gx.shape
    u, v = np.empty((h, w), np.float32), np.empty((h, w), np.float32)
    _check(lib().of_lk_from_gradients_f32(_ptr(gx), _ptr(gy), _ptr(gt), _ptr(u), _ptr(v), h, w, _window(window_size)))
    return u, v


def lk_single_scale(frame_prev, frame_curr, window_size: int = 5, mode: Optional[int] = None):
    p, c = _frame(frame_prev, "frame_prev"), _frame(frame_curr, "frame_curr")
    _same_shape(p, c)
    h, w = p.shape
    u, v = np.empty((h, w), np.float32), np.empty((h, w), np.float32)
    m = default_mode() if mode is None else int(mode)
    _check(lib().of_lk_single_scale_f32(_ptr(p), _ptr(c), _ptr(u), _ptr(v), 1, h, w, _window(window_size), m))
    return u, v


def lk_single_scale_batch(prev, curr, window_size: int = 5, mode: Optional[int] = None, out=None):
    """[B, H, W] float32 stacks of independent frame pairs -> (u, v) stacks."""
    p = np.ascontiguousarray(prev, dtype=np.float32)
    c = np.ascontiguousarray(curr, dtype=np.float32)
    if p.ndim != 3 or p.shape != c.shape:
        raise ValueError("prev and curr must be [B, H, W] arrays of equal shape")
    b, h, w = p.shape
    if out is None:
        u, v = np.empty_like(p), np.empty_like(p)
    else:
        u, v = _out_pair(out, p.shape)
    m = default_mode() if mode is None else int(mode)
    _check(lib().of_lk_single_scale_f32(_ptr(p), _ptr(c), _ptr(u), _ptr(v), b, h, w, _window(window_size), m))
    return u, v


def pyramid_down(image, scale_factor: float = 0.5):
    """One coarser level: gaussian_filter(sigma = 1/scale_factor) + bilinear decimation."""
    src = _frame(image, "image")
    h, w = src.shape
    oh, ow = int(h * scale_factor), int(w * scale_factor)
    if oh < 1 or ow < 1:
        raise ValueError("image too small to downsample")
    wts = gaussian_weights(1.0 / scale_factor)
    dst = np.empty((oh, ow), np.float32)
    _check(lib().of_pyramid_down_f32(_ptr(src), _ptr(dst), h, w, oh, ow, _ptr(wts), (len(wts) - 1) // 2))
    return dst


def warp(image, flow_u, flow_v):
    img, fu, fv = _frame(image, "image"), _frame(flow_u, "flow_u"), _frame(flow_v, "flow_v")
    _same_shape(img, fu, fv)
    h, w = img.shape
    out = np.empty((h, w), np.float32)
    _check(lib().of_warp_f32(_ptr(img), _ptr(fu), _ptr(fv), _ptr(out), h, w))
    return out


def upsample_flow(flow_u, flow_v, target_shape: Tuple[int, int]):
    cu, cv = _frame(flow_u, "flow_u"), _frame(flow_v, "flow_v")
    _same_shape(cu, cv)
    ch, cw = cu.shape
    th, tw = int(target_shape[0]), int(target_shape[1])
    u, v = np.empty((th, tw), np.float32), np.empty((th, tw), np.float32)
    _check(lib().of_upsample_flow_f32(_ptr(cu), _ptr(cv), _ptr(u), _ptr(v), ch, cw, th, tw))
    return u, v


def lk_pyramidal_batch(
    prev,
    curr,
    num_levels: int = 3,
    window_size: int = 5,
    num_iterations: int = 3,
    mode: Optional[int] = None,
    scale_factor: float = 0.5,
    return_trace: bool = False,
    out=None,
):
    """[B, H, W] stacks -> (u, v) stacks; optional trace = (iters_executed [B, L],
    residuals [B, L, I, 2]) with level 0 = coarsest, like the reference's loop index.
    out = (u, v): caller-supplied result stacks (e.g. pinned, so the library's copies overlap its kernels)."""
    p = np.ascontiguousarray(prev, dtype=np.float32)
    c = np.ascontiguousarray(curr, dtype=np.float32)
    if p.ndim != 3 or p.shape != c.shape:
        raise ValueError("prev and curr must be [B, H, W] arrays of equal shape")
    if scale_factor != 0.5:
        raise ValueError("only scale_factor = 0.5 (the reference's value) is supported")
    b, h, w = p.shape
    levels, iters = int(num_levels), int(num_iterations)
    u, v = (np.empty_like(p), np.empty_like(p)) if out is None else _out_pair(out, p.shape)
    wts = gaussian_weights(1.0 / scale_factor)
    it_exec = np.zeros((b, max(levels, 1)), np.int32)
    resid = np.zeros((b, max(levels, 1), max(iters, 1), 2), np.float32)
    m = default_mode() if mode is None else int(mode)
    _check(
        lib().of_lk_pyramidal_f32(
            _ptr(p), _ptr(c), _ptr(u), _ptr(v), b, h, w, levels, _window(window_size), iters, m,
            _ptr(wts), (len(wts) - 1) // 2, _ptr(it_exec), _ptr(resid),
        )
    )
    if return_trace:
        return u, v, (it_exec, resid)
    return u, v


def lk_pyramidal(frame_prev, frame_curr, num_levels=3, window_size=5, num_iterations=3, mode=None, return_trace=False):
    p, c = _frame(frame_prev, "frame_prev"), _frame(frame_curr, "frame_curr")
    _same_shape(p, c)
    res = lk_pyramidal_batch(p[None], c[None], num_levels, window_size, num_iterations, mode, 0.5, return_trace)
    if return_trace:
        return res[0][0], res[1][0], (res[2][0][0], res[2][1][0])
    return res[0][0], res[1][0]


def lk_single_scale_fx(prev_u8, curr_u8, mirror_avg_quirk: bool = True, out=None):
    """uint8 frame pair(s) -> int16 S8.7 flow (value / 128 = pixels), RTL integer datapath.
    out = (u, v): caller-supplied int16 result arrays of the frames' shape (e.g. pinned)."""
    p = np.ascontiguousarray(prev_u8, dtype=np.uint8)
    c = np.ascontiguousarray(curr_u8, dtype=np.uint8)
    if p.shape != c.shape or p.ndim not in (2, 3):
        raise ValueError("prev and curr must be uint8 arrays of equal shape, [H, W] or [B, H, W]")
    b = 1 if p.ndim == 2 else p.shape[0]
    h, w = p.shape[-2:]
    u, v = (np.empty(p.shape, np.int16), np.empty(p.shape, np.int16)) if out is None else _out_pair(out, p.shape, np.int16)
    flags = FX_MIRROR_AVG_QUIRK if mirror_avg_quirk else 0
    _check(lib().of_lk_single_scale_fx(_ptr(p), _ptr(c), _ptr(u), _ptr(v), b, h, w, flags))
    return u, v


# --------------------------------------------------------------------------------------
# pinned host memory and device-pointer API
# --------------------------------------------------------------------------------------
class PinnedArray:
    """A NumPy array whose storage is page-locked host memory (cudaHostAlloc), so the
    library's chunked H2D / D2H copies overlap with its kernels."""

    def __init__(self, shape: Sequence[int], dtype=np.float32):
        self.nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
        p = C.c_void_p()
        _check(lib().of_host_alloc_pinned(C.byref(p), self.nbytes))
        self._ptr = p
        buf = (C.c_char * self.nbytes).from_address(p.value)
        self.array = np.frombuffer(buf, dtype=dtype).reshape(shape)

    def free(self) -> None:
        if self._ptr is not None:
            self.array = None
            lib().of_host_free_pinned(self._ptr)
            self._ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def lk_single_scale_dev(prev_ptr, curr_ptr, u_ptr, v_ptr, batch, height, width, window_size=5, mode=MODE_FAST, stream=0):
    _check(
        lib().of_lk_single_scale_f32_dev(
            prev_ptr, curr_ptr, u_ptr, v_ptr, batch, height, width, _window(window_size), mode, stream
        )
    )


def lk_pyramidal_workspace_bytes(batch, height, width, levels, iterations) -> int:
    return int(lib().of_lk_pyramidal_workspace_bytes(batch, height, width, levels, iterations))


def lk_pyramidal_dev(
    prev_ptr, curr_ptr, u_ptr, v_ptr, batch, height, width, levels, window_size, iterations, mode,
    workspace_ptr, workspace_bytes, iters_ptr=None, resid_ptr=None, stream=0, weights: Optional[np.ndarray] = None,
):
    wts = gaussian_weights(2.0) if weights is None else np.ascontiguousarray(weights, dtype=np.float64)
    _check(
        lib().of_lk_pyramidal_f32_dev(
            prev_ptr, curr_ptr, u_ptr, v_ptr, batch, height, width, levels, _window(window_size), iterations, mode,
            _ptr(wts), (len(wts) - 1) // 2, workspace_ptr, workspace_bytes, iters_ptr, resid_ptr, stream,
        )
    )


def pyramid_down_dev(src_ptr, dst_ptr, batch, height, width, out_height, out_width, stream=0, sigma: float = 2.0,
                     row_lo: int = 0, row_hi: Optional[int] = None, mode: int = MODE_EXACT):
    wts = gaussian_weights(sigma)
    _check(
        lib().of_pyramid_down_f32_dev(
            src_ptr, dst_ptr, batch, height, width, out_height, out_width, _ptr(wts), (len(wts) - 1) // 2,
            row_lo, out_height if row_hi is None else row_hi, int(mode), stream,
        )
    )


def lk_refine_pingpong_dev(
    prev_ptr, curr_ptr, f0u, f0v, f1u, f1v, sel_ptr, done_ptr, batch, height, width, window_size, mode,
    row_lo, row_hi, own_lo, own_hi, sums_ptr, workspace_ptr, workspace_bytes, stream=0,
):
    _check(
        lib().of_lk_refine_pingpong_f32_dev(
            prev_ptr, curr_ptr, f0u, f0v, f1u, f1v, sel_ptr, done_ptr, batch, height, width, _window(window_size), mode,
            row_lo, row_hi, own_lo, own_hi, sums_ptr, workspace_ptr, workspace_bytes, stream,
        )
    )


def lk_convergence_update_dev(sums_ptr, batch, n_pixels, sel_ptr, done_ptr, iters_ptr, resid_ptr, max_iterations,
                              iteration, stream=0):
    _check(
        lib().of_lk_convergence_update_dev(
            sums_ptr, batch, float(n_pixels), sel_ptr, done_ptr, iters_ptr, resid_ptr, max_iterations, iteration, stream
        )
    )


def upsample_flow_dev(cu_ptr, cv_ptr, u_ptr, v_ptr, batch, ch, cw, th, tw, row_lo=0, row_hi=None, stream=0):
    _check(
        lib().of_upsample_flow_f32_dev(
            cu_ptr, cv_ptr, u_ptr, v_ptr, batch, ch, cw, th, tw, row_lo, th if row_hi is None else row_hi, stream
        )
    )


def lk_refine_workspace_bytes(batch, height, width) -> int:
    return int(lib().of_lk_refine_workspace_bytes(batch, height, width))


def lk_refine_dev(
    prev_ptr, curr_ptr, fin_u_ptr, fin_v_ptr, fout_u_ptr, fout_v_ptr, batch, height, width, window_size, mode,
    row_lo, row_hi, own_lo, own_hi, sums_ptr, workspace_ptr, workspace_bytes, stream=0,
):
    _check(
        lib().of_lk_refine_f32_dev(
            prev_ptr, curr_ptr, fin_u_ptr, fin_v_ptr, fout_u_ptr, fout_v_ptr, batch, height, width,
            _window(window_size), mode, row_lo, row_hi, own_lo, own_hi, sums_ptr, workspace_ptr, workspace_bytes, stream,
        )
    )


def lk_single_scale_fx_dev(prev_ptr, curr_ptr, u_ptr, v_ptr, batch, height, width, mirror_avg_quirk=True, stream=0):
    flags = FX_MIRROR_AVG_QUIRK if mirror_avg_quirk else 0
    _check(lib().of_lk_single_scale_fx_dev(prev_ptr, curr_ptr, u_ptr, v_ptr, batch, height, width, flags, stream))


def lk_single_scale_u8_batch(prev_u8, curr_u8, window_size: int = 5, mode: Optional[int] = None, out=None):
    """lucas_kanade_single_scale on [B, H, W] uint8 frame stacks (the on-disk value range): the same
    float32 flow as widening the frames first, with the widening done in registers on the GPU."""
    p = np.ascontiguousarray(prev_u8)
    c = np.ascontiguousarray(curr_u8)
    if p.dtype != np.uint8 or c.dtype != np.uint8:
        raise ValueError("frames must be uint8")
    if p.ndim != 3 or p.shape != c.shape:
        raise ValueError("prev and curr must be [B, H, W] arrays of equal shape")
    b, h, w = p.shape
    u, v = _out_pair(out, (b, h, w)) if out is not None else (np.empty((b, h, w), np.float32), np.empty((b, h, w), np.float32))
    m = default_mode() if mode is None else int(mode)
    _check(lib().of_lk_single_scale_u8(_ptr(p), _ptr(c), _ptr(u), _ptr(v), b, h, w, _window(window_size), m))
    return u, v


def lk_single_scale_u8_dev(prev_ptr, curr_ptr, u_ptr, v_ptr, batch, height, width, window_size=5, stream=0):
    _check(lib().of_lk_single_scale_u8_dev(prev_ptr, curr_ptr, u_ptr, v_ptr, batch, height, width, _window(window_size), stream))


def load_frame_u8(path, height: int, width: int) -> np.ndarray:
    """frame_XX.bin (raw bytes) or frame_XX.mem (one hex byte per line) -> uint8 [H, W]."""
    out = np.empty((int(height), int(width)), np.uint8)
    _check(lib().of_load_frame_u8(str(path).encode(), _ptr(out), int(height), int(width)))
    return out


def export_flow_txt(path, u, v, test_region=None) -> None:
    """export_flow_field_txt (lucas_kanade_reference.py:78-103) written by the native library.
    float32 flow -> the Python reference's header; int16 S8.7 flow (fixed-point mode) -> the RTL
    testbench's header.  test_region: dict with x_min / x_max / y_min / y_max, or None."""
    uu, vv = np.ascontiguousarray(u), np.ascontiguousarray(v)
    if uu.ndim != 2 or uu.shape != vv.shape or uu.dtype != vv.dtype:
        raise ValueError("u and v must be 2-D arrays of equal shape and dtype")
    h, w = uu.shape
    r = test_region or {"x_min": -1, "x_max": -1, "y_min": -1, "y_max": -1}
    args = (str(path).encode(), _ptr(uu), _ptr(vv), h, w, int(r["x_min"]), int(r["x_max"]), int(r["y_min"]), int(r["y_max"]))
    if uu.dtype == np.float32:
        _check(lib().of_export_flow_txt(*args))
    elif uu.dtype == np.int16:
        _check(lib().of_export_flow_fx_txt(*args))
    else:
        raise ValueError("flow must be float32 (pixels) or int16 (S8.7)")


def apply_motion_u8_batch(frames_u8, dx, dy, cval: float = 128.0) -> np.ndarray:
    """apply_motion (generate_test_frames_natural.py:67-73) for [B, H, W] (or [H, W]) uint8 frames with
    one (dx, dy) per frame: scipy.ndimage.shift(..., order=1, mode="constant", cval) bit for bit."""
    f = np.ascontiguousarray(frames_u8)
    if f.dtype != np.uint8:
        raise ValueError("frames must be uint8")
    single = f.ndim == 2
    if single:
        f = f[None]
    if f.ndim != 3:
        raise ValueError("frames must be [B, H, W] or [H, W]")
    b, h, w = f.shape
    xs = np.ascontiguousarray(np.broadcast_to(np.asarray(dx, dtype=np.float64), (b,)))
    ys = np.ascontiguousarray(np.broadcast_to(np.asarray(dy, dtype=np.float64), (b,)))
    out = np.empty_like(f)
    _check(lib().of_apply_motion_u8(_ptr(f), _ptr(out), b, h, w, _ptr(xs), _ptr(ys), float(cval)))
    return out[0] if single else out


def apply_motion_u8_dev(frames_ptr, out_ptr, batch, height, width, dx_ptr, dy_ptr, cval: float = 128.0, stream=0):
    _check(lib().of_apply_motion_u8_dev(frames_ptr, out_ptr, batch, height, width, dx_ptr, dy_ptr, float(cval), stream))


def motion_matrix(width: int, height: int, dx: float = 0.0, dy: float = 0.0, rotation: float = 0.0, scale: float = 1.0):
    """The 2x3 matrix apply_motion_opencv builds (generate_test_suite.py:183-190): cv2.getRotationMatrix2D
    about the frame centre (libm cos / sin of angle * (pi / 180)), then the translation."""
    import math

    a = rotation * (math.pi / 180.0)
    alpha, beta = math.cos(a) * scale, math.sin(a) * scale
    cx, cy = width / 2.0, height / 2.0
    return np.array([[alpha, beta, (1 - alpha) * cx - beta * cy + dx], [-beta, alpha, beta * cx + (1 - alpha) * cy + dy]],
                    np.float64)


def warp_affine_u8_batch(frames_u8, matrices, cval: int = 128) -> np.ndarray:
    """cv2.warpAffine(INTER_LINEAR, BORDER_CONSTANT, cval) for [B, H, W] (or [H, W]) uint8 frames with one
    2x3 forward matrix per frame; bit-identical to OpenCV (apply_motion_opencv of generate_test_suite.py)."""
    f = np.ascontiguousarray(frames_u8)
    if f.dtype != np.uint8:
        raise ValueError("frames must be uint8")
    single = f.ndim == 2
    if single:
        f = f[None]
    if f.ndim != 3:
        raise ValueError("frames must be [B, H, W] or [H, W]")
    b, h, w = f.shape
    m = np.ascontiguousarray(np.broadcast_to(np.asarray(matrices, dtype=np.float64).reshape(-1, 2, 3), (b, 2, 3)))
    out = np.empty_like(f)
    _check(lib().of_warp_affine_u8(_ptr(f), _ptr(out), b, h, w, _ptr(m), int(cval)))
    return out[0] if single else out


METRIC_NAMES = ("mae_u", "mae_v", "rmse", "epe", "aae")


def verifier_test_region(shape, pattern_type: str, center_crop_size: int = 80):
    """(y0, y1, x0, x1) of the verifier's test-region mask (optical_flow_verifier.py:96-138):
    rotation / zoom / combined patterns -> central crop, translations -> frame minus a 10 px border."""
    h, w = int(shape[0]), int(shape[1])
    if "rotate" in pattern_type or "zoom" in pattern_type:
        cy, cx, half = h // 2, w // 2, int(center_crop_size) // 2
        return cy - half, cy + half, cx - half, cx + half
    return 10, h - 10, 10, w - 10


def flow_metrics_batch(u, v, u_true, v_true, region=None):
    """compute_all_metrics (flow_metrics.py:166-201) for [B, H, W] flow stacks on the GPU.
    u_true / v_true: one constant ground-truth flow per pair.  region = (y0, y1, x0, x1) or None
    for the whole frame.  Returns a list of dicts with the reference's keys."""
    uu = np.ascontiguousarray(u, dtype=np.float32)
    vv = np.ascontiguousarray(v, dtype=np.float32)
    if uu.ndim == 2:
        uu, vv = uu[None], vv[None]
    if uu.ndim != 3 or uu.shape != vv.shape:
        raise ValueError("u and v must be [B, H, W] (or [H, W]) arrays of equal shape")
    b, h, w = uu.shape
    ut = np.ascontiguousarray(np.broadcast_to(np.asarray(u_true, dtype=np.float32), (b,)))
    vt = np.ascontiguousarray(np.broadcast_to(np.asarray(v_true, dtype=np.float32), (b,)))
    y0, y1, x0, x1 = (0, h, 0, w) if region is None else (int(r) for r in region)
    out = np.zeros((b, 5), dtype=np.float64)
    _check(lib().of_flow_metrics_f32(_ptr(uu), _ptr(vv), _ptr(ut), _ptr(vt), b, h, w, y0, y1, x0, x1, _ptr(out)))
    return [dict(zip(METRIC_NAMES, (float(x) for x in row))) for row in out]


def flow_metrics_workspace_bytes(batch, region_height, region_width) -> int:
    return int(lib().of_flow_metrics_workspace_bytes(batch, region_height, region_width))


def flow_metrics_dev(u_ptr, v_ptr, u_true_ptr, v_true_ptr, batch, height, width, region, metrics_ptr, workspace_ptr,
                     workspace_bytes, stream=0):
    y0, y1, x0, x1 = (int(r) for r in region)
    _check(
        lib().of_flow_metrics_f32_dev(
            u_ptr, v_ptr, u_true_ptr, v_true_ptr, batch, height, width, y0, y1, x0, x1, metrics_ptr, workspace_ptr,
            workspace_bytes, stream,
        )
    )


IPC_HANDLE_BYTES = 64


class RowbandContext:
    """Native row-band (multi-GPU) pyramidal LK for ONE frame pair split over `world` ranks.

    Owns this rank's peer-memory arena.  Connect the ranks with `ipc_handle()` /
    `open_peers_ipc(handles)` (processes; the handles travel over any channel) or
    `set_peers(arena_pointers)` (threads of one process), then call `run(prev_ptr, curr_ptr)` on
    every rank: one C call enqueues the whole coarse-to-fine computation, and the ranks exchange
    pyramid rows, flow rows and residual sums through peer stores inside the kernels.
    """

    def __init__(self, rank, world, height, width, num_levels=3, window_size=5, num_iterations=3, mode=MODE_FAST,
                 sigma: float = 2.0):
        self.rank, self.world, self.height, self.width = int(rank), int(world), int(height), int(width)
        self.levels, self.iterations = int(num_levels), int(num_iterations)
        wts = gaussian_weights(sigma)
        self._ctx = _vp()
        _check(
            lib().of_rowband_create(
                C.byref(self._ctx), self.rank, self.world, self.height, self.width, self.levels, _window(window_size),
                self.iterations, mode, _ptr(wts), (len(wts) - 1) // 2,
            )
        )

    @property
    def arena_ptr(self) -> int:
        return int(lib().of_rowband_arena(self._ctx) or 0)

    @property
    def arena_bytes(self) -> int:
        return int(lib().of_rowband_arena_bytes(self._ctx))

    def ipc_handle(self) -> bytes:
        buf = C.create_string_buffer(IPC_HANDLE_BYTES)
        _check(lib().of_rowband_ipc_handle(self._ctx, buf))
        return buf.raw

    def open_peers_ipc(self, handles) -> None:
        blob = b"".join(handles)
        if len(blob) != self.world * IPC_HANDLE_BYTES:
            raise ValueError("need one 64-byte handle per rank, in rank order")
        _check(lib().of_rowband_open_peers_ipc(self._ctx, C.c_char_p(blob)))

    def set_peers(self, arena_ptrs) -> None:
        if len(arena_ptrs) != self.world:
            raise ValueError("need one arena pointer per rank, in rank order")
        arr = (_vp * self.world)(*[_vp(int(p)) for p in arena_ptrs])
        _check(lib().of_rowband_set_peers(self._ctx, arr))

    def set_replicate_pixels(self, pixels: int) -> None:
        """Levels of at most `pixels` pixels are computed whole on every rank (default 600000, 0 = never)."""
        _check(lib().of_rowband_set_replicate_pixels(self._ctx, int(pixels)))

    def set_timeout_ms(self, milliseconds: int) -> None:
        """Peer wait time-out (default 4000 ms); the ranks must enter run() within it of each other."""
        _check(lib().of_rowband_set_timeout_ms(self._ctx, int(milliseconds)))

    def run(self, prev_ptr, curr_ptr, u_ptr=None, v_ptr=None, stream=0) -> None:
        _check(lib().of_rowband_run(self._ctx, prev_ptr, curr_ptr, u_ptr, v_ptr, stream))

    def result_ptrs(self):
        u, v = _vp(), _vp()
        _check(lib().of_rowband_result(self._ctx, C.byref(u), C.byref(v)))
        return int(u.value), int(v.value)

    def trace(self, stream=0):
        """(iters_executed[levels], residuals[levels, iterations, 2], 0) after waiting for `stream`.
        Raises OFBackendError if a peer rank did not answer within the time-out during the run (the flow of
        that run is garbage); the next run starts clean."""
        iters = np.zeros(self.levels, dtype=np.int32)
        resid = np.zeros((self.levels, max(self.iterations, 1), 2), dtype=np.float32)
        err = C.c_int(0)
        _check(lib().of_rowband_trace(self._ctx, _ptr(iters), _ptr(resid) if self.iterations > 0 else None, C.byref(err), stream))
        return iters, resid[:, : self.iterations], int(err.value)

    def status(self, stream=0) -> None:
        """Waits for `stream`; raises OFBackendError if the run saw a peer time-out."""
        _check(lib().of_rowband_status(self._ctx, stream))

    def close(self) -> None:
        if self._ctx:
            lib().of_rowband_destroy(self._ctx)
            self._ctx = _vp()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
