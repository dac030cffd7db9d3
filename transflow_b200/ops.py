"""Thin, typed wrappers over the C ABI: torch CUDA tensors in, torch CUDA tensors out.

Everything here runs on the current CUDA device and the current torch stream.  No function
has a CPU path; CPU tensors raise.
"""
import ctypes as C
import os

import torch

from . import _lib
from ._lib import check, ptr, stream_ptr


#: solve kernel variant used when none is requested: 8 = half-buffer kernel (configuration 24) on the large pyramid
#: levels, rolling-tile kernel on the small ones; 3 rolling tile everywhere; 4-7, 17, 23, 24 half-buffer configurations;
#: 9-16 staged experiments; 18-22 ring kernels (R1 taps from a TMA-staged rolling row ring in shared memory); 0 / 2 fused
#: streaming (float / double sums in shared memory); 1 unfused reference kernels
DEFAULT_FB_VARIANT = 8


def _cuda(t: torch.Tensor, dtype, name: str) -> torch.Tensor:
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor (transflow_b200 has no CPU fallback)")
    if t.dtype != dtype:
        raise ValueError(f"{name} must be {dtype}, got {t.dtype}")
    return t if t.is_contiguous() else t.contiguous()


def gray_from_bgr(bgr: torch.Tensor, out: torch.Tensor | None = None) -> torch.Tensor:
    """``cv2.cvtColor(frame, COLOR_BGR2GRAY)`` (flow/sources/cv.py:465), bit-exact."""
    bgr = _cuda(bgr, torch.uint8, "bgr")
    if bgr.ndim != 3 or bgr.shape[2] != 3:
        raise ValueError(f"bgr must be (H, W, 3), got {tuple(bgr.shape)}")
    h, w = bgr.shape[:2]
    if out is None:
        out = torch.empty((h, w), dtype=torch.uint8, device=bgr.device)
    check(_lib.load().tf_gray_from_bgr(ptr(bgr), ptr(out), h, w, stream_ptr()))
    return out


def resize_nearest_bgr(bgr: torch.Tensor, height: int, width: int, out: torch.Tensor | None = None) -> torch.Tensor:
    """``cv2.resize(frame, (width, height), interpolation=cv2.INTER_NEAREST)`` (flow/sources/cv.py:464), bit-exact."""
    bgr = _cuda(bgr, torch.uint8, "bgr")
    if bgr.ndim != 3 or bgr.shape[2] != 3:
        raise ValueError(f"bgr must be (H, W, 3), got {tuple(bgr.shape)}")
    if out is None:
        out = torch.empty((int(height), int(width), 3), dtype=torch.uint8, device=bgr.device)
    check(_lib.load().tf_resize_nearest_bgr(ptr(bgr), int(bgr.shape[0]), int(bgr.shape[1]), ptr(out), int(height),
                                            int(width), stream_ptr()))
    return out


class Farneback:
    """Device Farneback flow with the parameters of ``CvFlowConfig.fb_*`` (cv.py:275-281)."""

    def __init__(self, height, width, pyr_scale=0.5, levels=3, winsize=15, iterations=3, poly_n=5,
                 poly_sigma=1.2, flags=0, r_fp16=None, variant=None, debug=False, lanes=1):
        self.lib = _lib.load()
        self.h, self.w = int(height), int(width)
        if variant is None:
            variant = int(os.environ.get("TFB200_FB_VARIANT", DEFAULT_FB_VARIANT))
        if r_fp16 is None:      # half-precision STORAGE of the polynomial expansion (compute stays fp32)
            r_fp16 = bool(int(os.environ.get("TFB200_R_FP16", "0")))
        self.variant = int(variant)
        self.handle = C.c_void_p()
        check(self.lib.tf_farneback_create(C.byref(self.handle), self.h, self.w, float(pyr_scale), int(levels),
                                           int(winsize), int(iterations), int(poly_n), float(poly_sigma),
                                           int(flags), int(bool(r_fp16))))
        if lanes > 1:   # two pairs in flight: third frame slot + second solve lane allocated now, not in the frame loop
            check(self.lib.tf_farneback_reserve(self.handle, 3, int(lanes)))
        if debug:       # keep every pyramid image readable through debug_read
            check(self.lib.tf_farneback_set_debug(self.handle, 1))

    def close(self):
        handle = getattr(self, "handle", None)
        if handle is not None and handle.value:
            self.handle = None          # (module globals may already be gone at interpreter shutdown)
            self.lib.tf_farneback_destroy(handle)

    __del__ = close

    def _gray(self, g, name):
        g = _cuda(g, torch.uint8, name)
        if tuple(g.shape) != (self.h, self.w):
            raise ValueError(f"{name} must be ({self.h}, {self.w}), got {tuple(g.shape)}")
        return g

    def prepare(self, slot: int, gray: torch.Tensor):
        check(self.lib.tf_farneback_prepare(self.handle, int(slot), ptr(self._gray(gray, "gray")), stream_ptr()))

    def solve(self, slot_left: int, slot_right: int, out=None, clip=False) -> torch.Tensor:
        if out is None:
            out = torch.empty((self.h, self.w, 2), dtype=torch.float32, device="cuda")
        check(self.lib.tf_farneback_solve(self.handle, int(slot_left), int(slot_right), ptr(out), self.variant,
                                          int(bool(clip)), stream_ptr()))
        return out

    def step(self, new_slot: int, gray: torch.Tensor, slot_left: int, slot_right: int, out=None,
             clip=False, lane: int = 0) -> torch.Tensor:
        """prepare(new_slot, gray) overlapped with solve(slot_left, slot_right) (streaming sources).
        ``lane`` (0 or 1) selects one of two sets of per-level flow buffers, so that two pairs can be in
        flight on two streams (frame t in slot t % 3, pair t on lane t % 2)."""
        if out is None:
            out = torch.empty((self.h, self.w, 2), dtype=torch.float32, device="cuda")
        check(self.lib.tf_farneback_step_lane(self.handle, int(lane), int(new_slot), ptr(self._gray(gray, "gray")),
                                              int(slot_left), int(slot_right), ptr(out), self.variant,
                                              int(bool(clip)), stream_ptr()))
        return out

    def __call__(self, left: torch.Tensor, right: torch.Tensor, out=None) -> torch.Tensor:
        self.prepare(0, left)
        self.prepare(1, right)
        return self.solve(0, 1, out)

    @property
    def level_sizes(self):
        n = self.lib.tf_farneback_num_levels(self.handle)
        out = []
        for i in range(n):
            w, h = C.c_int(), C.c_int()
            check(self.lib.tf_farneback_level_size(self.handle, i, C.byref(w), C.byref(h)))
            out.append((h.value, w.value))
        return out

    def debug_read(self, slot: int, level_index: int, what: int) -> torch.Tensor:
        h, w = self.level_sizes[level_index]
        shape = {0: (h, w), 1: (5, h, w), 2: (h, w, 2)}[what]
        out = torch.empty(shape, dtype=torch.float32, device="cuda")
        check(self.lib.tf_farneback_debug_read(self.handle, slot, level_index, what, ptr(out), stream_ptr()))
        return out

    def algorithmic_bytes(self, reuse_r=True) -> float:
        return float(self.lib.tf_farneback_algorithmic_bytes(self.handle, int(bool(reuse_r))))


class HornSchunck:
    """Device Horn-Schunck (flow/methods/horn_schunck.py:9-45)."""

    def __init__(self, height, width):
        self.lib = _lib.load()
        self.h, self.w = int(height), int(width)
        self.handle = C.c_void_p()
        check(self.lib.tf_hs_create(C.byref(self.handle), self.h, self.w))
        self.last_sweeps = None
        self.track_sweeps = False

    def close(self):
        handle = getattr(self, "handle", None)
        if handle is not None and handle.value:
            self.handle = None          # (module globals may already be gone at interpreter shutdown)
            self.lib.tf_hs_destroy(handle)

    __del__ = close

    def __call__(self, left, right, flow=None, alpha=1, max_iters=3, decay=0, delta=1, out=None, clip=False):
        left = _cuda(left, torch.uint8, "left")
        right = _cuda(right, torch.uint8, "right")
        if tuple(left.shape) != (self.h, self.w) or tuple(right.shape) != (self.h, self.w):
            raise ValueError(f"frames must be ({self.h}, {self.w})")
        if flow is not None:
            flow = _cuda(flow, torch.float32, "flow")
        if out is None:
            out = torch.empty((self.h, self.w, 2), dtype=torch.float32, device="cuda")
        sweeps = C.c_int(0)
        # asking for the sweep count makes the call wait for the device; off by default
        check(self.lib.tf_hs_run(self.handle, ptr(left), ptr(right), ptr(flow), float(alpha), int(max_iters),
                                 float(decay), -1.0 if delta is None else float(delta), ptr(out),
                                 int(bool(clip)), C.byref(sweeps) if self.track_sweeps else None, stream_ptr()))
        self.last_sweeps = sweeps.value if self.track_sweeps else None
        return out


class LucasKanade:
    """Device pyramidal Lucas-Kanade on a dense grid (flow/methods/lukas_kanade.py:9-36)."""

    def __init__(self, height, width, win_size=15, max_level=2, step=1):
        self.lib = _lib.load()
        self.h, self.w = int(height), int(width)
        self.handle = C.c_void_p()
        check(self.lib.tf_lk_create(C.byref(self.handle), self.h, self.w, int(win_size), int(max_level), int(step)))

    def close(self):
        handle = getattr(self, "handle", None)
        if handle is not None and handle.value:
            self.handle = None          # (module globals may already be gone at interpreter shutdown)
            self.lib.tf_lk_destroy(handle)

    __del__ = close

    def __call__(self, left, right, out=None, clip=False):
        left = _cuda(left, torch.uint8, "left")
        right = _cuda(right, torch.uint8, "right")
        if tuple(left.shape) != (self.h, self.w) or tuple(right.shape) != (self.h, self.w):
            raise ValueError(f"frames must be ({self.h}, {self.w})")
        if out is None:
            out = torch.empty((self.h, self.w, 2), dtype=torch.float32, device="cuda")
        check(self.lib.tf_lk_run(self.handle, ptr(left), ptr(right), ptr(out), int(bool(clip)), stream_ptr()))
        return out


FLOW_OP_KINDS = {"scale": 0, "threshold": 1, "clip": 2}
MAX_FLOW_OPS = 8


def _pack_flow_ops(ops):
    """[(kind name, scalar), ...] -> (ctypes array | None, count).  A NumPy floating scalar keeps NumPy's
    "strong" float64 arithmetic, a Python number is weak (float32 next to the float32 flow)."""
    import numpy as np
    ops = list(ops or ())
    if len(ops) > MAX_FLOW_OPS:
        raise ValueError(f"at most {MAX_FLOW_OPS} elementwise flow filters per call, got {len(ops)}")
    if not ops:
        return None, 0
    arr = (_lib.FlowOpStruct * len(ops))()
    for slot, (kind, value) in zip(arr, ops):
        slot.kind = FLOW_OP_KINDS[kind]
        slot.strong = int(isinstance(value, np.floating) and not isinstance(value, np.float32))
        slot.value = float(np.float32(value)) if isinstance(value, np.float32) else float(value)
    return arr, len(ops)


class ForwardClaims:
    """A FORWARD flow in the form the scatter pass of ``post_process`` leaves it (source.py:349-360): an int32 plane in
    which every moved pixel claimed its target (``numpy.put``: the last source in raster order wins).  The flow the
    reference returns is "claimant - own position"; the move-reference compositor layer only needs the claimant, so
    ``Compositor.step`` consumes the plane directly (``tf_layer_update_claims``) and that flow is never written to or
    read from HBM.  Any other consumer calls ``tensor()``, which forms the flow (once) exactly as ``PostProcess`` would
    have.  The plane belongs to the ``PostProcess`` that made it and is recycled: use a claims object before asking
    the source for ``PostProcess.CLAIM_PLANES`` more flows."""

    def __init__(self, post, plane):
        self._post, self._plane, self._flow, self._state = post, plane, None, "claims"
        self.shape = (post.h, post.w, 2)

    @classmethod
    def from_plane(cls, post, plane: torch.Tensor):
        """A claims object over a claim plane that did not come from ``post``'s own ring: int32 (H, W) device memory
        filled by ``tf_flow_forward_claims`` somewhere else (another GPU's scatter pass, copied into a ring slot)."""
        if plane.dtype != torch.int32 or tuple(plane.shape) != (post.h, post.w) or not plane.is_contiguous():
            raise ValueError(f"a claim plane is a contiguous int32 ({post.h}, {post.w}) tensor")
        return cls(post, plane)

    @property
    def live(self) -> bool:
        """True while the claim plane still holds this flow (not consumed, not turned into a tensor, not recycled)."""
        return self._state == "claims"

    def take(self) -> torch.Tensor:
        """Hand the claim plane to a consumer that zeroes it (the compositor layer)."""
        if self._state != "claims":
            raise RuntimeError(f"this forward flow's claim plane is gone ({self._state})")
        self._state = "consumed by the compositor"
        return self._plane

    def tensor(self) -> torch.Tensor:
        """The flow as the (H, W, 2) float32 device tensor ``PostProcess.__call__`` returns."""
        if self._state == "tensor":
            return self._flow
        if self._state != "claims":
            raise RuntimeError(f"this forward flow's claim plane is gone ({self._state}): ask for the tensor before "
                               "the compositor consumes it, or set the source's output to 'device'")
        out = torch.empty(self.shape, dtype=torch.float32, device="cuda")
        check(self._post.lib.tf_flow_from_claims(ptr(self._plane), ptr(out), self._post.h, self._post.w, stream_ptr()))
        self._flow, self._state = out, "tensor"
        return out

    def _expire(self):
        if self._state == "claims":
            self._plane.zero_()
            self._state = "recycled before use"

    def cpu(self):
        return self.tensor().cpu()


class PostProcess:
    """``FlowSource.post_process`` (flow/sources/source.py:337-363) on a device flow: elementwise filters ->
    mask -> [convolution kernel] -> [forward: clip, round, scatter] -> clip."""

    def __init__(self, height, width, forward: bool, mask: torch.Tensor | None = None, kernel=None):
        self.lib = _lib.load()
        self.h, self.w, self.forward = int(height), int(width), bool(forward)
        self.mask = None if mask is None else _cuda(mask, torch.float32, "mask").reshape(self.h, self.w)
        self.owner = torch.zeros((self.h, self.w), dtype=torch.int32, device="cuda") if self.forward else None
        self.kernel = None
        if kernel is not None:      # scipy.signal.convolve2d promotes to float64 (source.py:345-347)
            self.kernel = torch.as_tensor(kernel, dtype=torch.float64).contiguous().cuda()
            if self.kernel.ndim != 2:
                raise ValueError(f"the flow kernel must be 2-D, got shape {tuple(self.kernel.shape)}")
            self._tmp = torch.empty((self.h, self.w, 2), dtype=torch.float32, device="cuda")
        self._claim_ring, self._claim_next = [], 0

    CLAIM_PLANES = 4

    def claims(self, flow: torch.Tensor, ops=None, plane: torch.Tensor | None = None) -> ForwardClaims:
        """Forward direction without a convolution kernel: filters + mask + clip + round + the scatter pass only; the
        result is a ``ForwardClaims`` (``flow`` itself is left untouched).  ``plane``: the caller's own all-zero int32
        (H, W) plane instead of one of this object's recycled ones."""
        if not self.forward or self.kernel is not None:
            raise ValueError("claims() is the forward direction's scatter pass (no convolution kernel)")
        flow = _cuda(flow, torch.float32, "flow")
        if tuple(flow.shape) != (self.h, self.w, 2):
            raise ValueError(f"flow must be ({self.h}, {self.w}, 2), got {tuple(flow.shape)}")
        if plane is not None:
            made = ForwardClaims.from_plane(self, plane)
            arr, n = _pack_flow_ops(ops)
            check(self.lib.tf_flow_forward_claims(ptr(flow), arr, n, ptr(self.mask), ptr(plane), self.h, self.w,
                                                  stream_ptr()))
            return made
        if len(self._claim_ring) < self.CLAIM_PLANES:
            self._claim_ring.append([torch.zeros((self.h, self.w), dtype=torch.int32, device="cuda"), None])
        slot = self._claim_ring[self._claim_next % len(self._claim_ring)]
        self._claim_next += 1
        if slot[1] is not None:
            slot[1]._expire()       # an unused claims object from CLAIM_PLANES flows ago: its plane is cleared
        arr, n = _pack_flow_ops(ops)
        check(self.lib.tf_flow_forward_claims(ptr(flow), arr, n, ptr(self.mask), ptr(slot[0]), self.h, self.w,
                                              stream_ptr()))
        slot[1] = ForwardClaims(self, slot[0])
        return slot[1]

    def __call__(self, flow: torch.Tensor, out: torch.Tensor | None = None, ops=None) -> torch.Tensor:
        """In place by default; ``out`` (same shape, may be peer memory) receives the result instead and ``flow``
        keeps its value.  ``ops``: [("scale" | "threshold" | "clip", scalar), ...] applied first, in order."""
        flow = _cuda(flow, torch.float32, "flow")
        if tuple(flow.shape) != (self.h, self.w, 2):
            raise ValueError(f"flow must be ({self.h}, {self.w}, 2), got {tuple(flow.shape)}")
        arr, n = _pack_flow_ops(ops)
        if self.kernel is not None:
            check(self.lib.tf_flow_filters(ptr(flow), arr, n, ptr(self.mask), ptr(self._tmp), self.h, self.w,
                                           stream_ptr()))
            kh, kw = self.kernel.shape
            dst = flow if out is None else out
            check(self.lib.tf_flow_convolve(ptr(self._tmp), ptr(self.kernel), kh, kw, int(self.forward), ptr(dst),
                                            self.h, self.w, stream_ptr()))
            check(self.lib.tf_flow_postprocess_ex(ptr(dst), None, 0, None, int(self.forward), ptr(self.owner),
                                                  None, self.h, self.w, stream_ptr()))
            return dst
        check(self.lib.tf_flow_postprocess_ex(ptr(flow), arr, n, ptr(self.mask), int(self.forward), ptr(self.owner),
                                              ptr(out), self.h, self.w, stream_ptr()))
        return flow if out is None else out

    @property
    def copies(self) -> bool:
        """True when the reference's ``post_process`` leaves its input alone after the filters (``numpy.multiply``
        by the mask and ``numpy.stack`` of the convolved channels make copies, source.py:342-348)."""
        return self.mask is not None or self.kernel is not None


MERGE_MODES = {"first": 0, "sum": 1, "average": 2, "difference": 3, "product": 4, "maskbin": 5, "masklin": 6,
               "absmax": 7}


def merge_flows(flows, mode: str, out: torch.Tensor | None = None) -> torch.Tensor:
    """``Pipeline.FLOW_MERGING_FUNCTIONS[mode](flows)`` (pipeline.py:149-158) on device flows."""
    if mode not in MERGE_MODES:
        raise KeyError(mode)
    flows = [_cuda(f, torch.float32, "flow") for f in flows]
    if not flows:
        raise ValueError("no flow to merge")
    h, w = flows[0].shape[:2]
    if any(tuple(f.shape) != (h, w, 2) for f in flows):
        raise ValueError("flows to merge must share one shape (H, W, 2)")
    if out is None:
        out = torch.empty_like(flows[0])
    arr = (C.c_void_p * len(flows))(*[f.data_ptr() for f in flows])
    check(_lib.load().tf_flow_merge(arr, len(flows), MERGE_MODES[mode], ptr(out), h, w, stream_ptr()))
    return out


def upscale_flow(flow: torch.Tensor, wf: int, hf: int) -> torch.Tensor:
    """``utils.upscale_array(flow, wf, hf)`` (utils.py:417-418) on a device flow."""
    flow = _cuda(flow, torch.float32, "flow")
    h, w = flow.shape[:2]
    out = torch.empty((h * int(hf), w * int(wf), 2), dtype=torch.float32, device=flow.device)
    check(_lib.load().tf_flow_upscale(ptr(flow), ptr(out), h, w, int(wf), int(hf), stream_ptr()))
    return out


RENDER_2D, RENDER_MAGNITUDE, RENDER_1D = 0, 1, 2


def render_flow(arr: torch.Tensor, mode: int, scale: float, colors, binary: bool = False,
                out: torch.Tensor | None = None) -> torch.Tensor:
    """``render2d`` / ``render1d`` (output/render.py:9-48) on a device array -> uint8 (H, W, 3).
    ``colors``: 4 (mode 0) or 2 (modes 1, 2) RGB triples in 0..255."""
    arr = _cuda(arr, torch.float32, "arr")
    want = 2 if mode in (RENDER_2D, RENDER_MAGNITUDE) else None
    if (want and (arr.ndim != 3 or arr.shape[2] != 2)) or (not want and arr.ndim not in (2, 3)):
        raise ValueError(f"render mode {mode}: unexpected array shape {tuple(arr.shape)}")
    if not want and arr.ndim == 3 and arr.shape[2] != 1:
        raise ValueError(f"render1d takes a scalar (H, W) array, got {tuple(arr.shape)}")
    h, w = arr.shape[:2]
    cols = (C.c_float * (3 * len(colors)))(*[float(v) for c in colors for v in c])
    if out is None:
        out = torch.empty((h, w, 3), dtype=torch.uint8, device=arr.device)
    check(_lib.load().tf_flow_render(ptr(arr), int(mode), float(scale), cols, len(colors), int(bool(binary)), ptr(out),
                                     h, w, stream_ptr()))
    return out
