"""Compositor layers backed by device state (drop-in for ``transflow/compositor/layers/*``).

Public surface kept from the reference (``layer.py:11-55``, ``data.py:6-17``):
``Layer(config, height, width, sources)``, ``set_sources``, ``update(flow)``,
``render() -> uint8 (H, W, 4)``, ``from_args`` string dispatch, and the state attributes other
code reads -- ``data`` (int32 (H, W, DEPTH)), ``rgba``, ``base``, ``INDEX_I/J/ALPHA/SOURCE``,
``DEPTH``, ``sources``.  ``data`` / ``rgba`` are materialised from HBM on access, and pickling goes
through the same NumPy arrays plus the mask arrays (the reference pickles its mask arrays too, so a
``random`` mask survives a checkpoint).  The pickled class paths differ from the reference's, so a
``.ckpt.zip`` is resumed by the side that wrote it.
"""
import ctypes as C
import os

import numpy as np
import torch

from ... import _lib
from ..._lib import LAYER_KINDS, RESET_MODES, LayerConfigStruct, PixmapStruct, check, ptr, stream_ptr
from ...config import LayerConfig
from ...ops import ForwardClaims
from ...utils import load_bool_mask, load_float_mask


def _to_device_u8(img):
    """NumPy or torch (H, W, C) uint8 -> contiguous CUDA tensor."""
    if isinstance(img, torch.Tensor):
        if not img.is_cuda:
            img = img.cuda(non_blocking=True)
        if img.dtype != torch.uint8:
            raise AssertionError("pixmap dtype must be uint8")
        return img.contiguous()
    if img.dtype != np.uint8:
        raise AssertionError("pixmap dtype must be uint8")
    return torch.from_numpy(np.ascontiguousarray(img)).cuda(non_blocking=True)


def flow_to_device(flow) -> torch.Tensor:
    if isinstance(flow, torch.Tensor):
        t = flow if flow.is_cuda else flow.cuda(non_blocking=True)
    else:
        t = torch.from_numpy(np.ascontiguousarray(flow, dtype=np.float32)).cuda(non_blocking=True)
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


class Layer:
    """Base class; concrete kinds differ only in ``KIND`` and the record layout constants."""

    KIND = None
    DEPTH = 4
    INDEX_I = 0
    INDEX_J = 1
    INDEX_ALPHA = 2
    INDEX_SOURCE = 3

    def __init__(self, config: LayerConfig, height: int, width: int, sources):
        if self.KIND is None:
            raise NotImplementedError()
        self.config = config
        self.height = int(height)
        self.width = int(width)
        self.sources = list(sources)
        #: "device" = counter-based Philox draws on the GPU (throughput mode);
        #: "numpy"  = numpy.random.random((H, W)) on the host each frame, uploaded -- the exact
        #:            draws the reference consumes (reference.py:59), for bit-exact parity.
        self.reset_rng = os.environ.get("TRANSFLOW_B200_RESET_RNG", "device")
        self.rng_seed = 0x5EED
        self.set_seed(None)
        self._lib = _lib.load()
        self._handle = C.c_void_p()
        self._create()
        self._upload_masks()
        if self.sources:
            self._push_sources()

    # -- construction ---------------------------------------------------------------------------
    def _config_struct(self) -> LayerConfigStruct:
        c = self.config
        if c.reset_mode not in RESET_MODES:
            raise ValueError(f"Unknown reset mode {c.reset_mode}")
        s = LayerConfigStruct()
        s.kind = LAYER_KINDS[self.KIND]
        s.reset_mode = RESET_MODES[c.reset_mode]
        for name in ("transparent_pixels_can_move", "pixels_can_move_to_empty_spot",
                     "pixels_can_move_to_filled_spot", "moving_pixels_leave_empty_spot", "reset_source",
                     "introduce_pixels_on_empty_spots", "introduce_pixels_on_filled_spots",
                     "introduce_moving_pixels", "introduce_unmoving_pixels", "introduce_once",
                     "introduce_on_all_filled_spots", "introduce_on_all_empty_spots"):
            setattr(s, name, int(bool(getattr(c, name))))
        s.reset_constant_step = float(np.float32(c.reset_constant_step))
        s.reset_random_factor = float(np.float32(c.reset_random_factor))
        s.reset_linear_factor = float(c.reset_linear_factor)
        return s

    def _create(self):
        cfg = self._config_struct()
        check(self._lib.tf_layer_create(C.byref(self._handle), self.height, self.width, C.byref(cfg)))

    def set_seed(self, seed):
        """Philox key of the device reset draws: one stream per (``config.seed``, layer index), so two layers with
        ``reset_mode=random`` draw different numbers and ``Config.seed`` reproduces a run."""
        base = 0 if seed is None else int(seed)
        index = int(getattr(self.config, "index", 0) or 0)
        self.rng_seed = (0x5EED + base * 0x9E3779B97F4A7C15 + index * 0xC2B2AE3D27D4EB4F) & 0xFFFFFFFFFFFFFFFF

    def _upload_masks(self, arrays=None):
        """``arrays``: the four mask arrays of a pickled layer (else they are built from the config strings)."""
        shape = (self.height, self.width)
        c = self.config
        if arrays is not None:
            self.mask_alpha, self.mask_src, self.mask_dst, self.reset_mask = arrays
        else:
            self.mask_alpha = load_float_mask(c.mask_alpha, shape, 1)
            self.mask_src = load_bool_mask(c.mask_src, shape, True)
            self.mask_dst = load_bool_mask(c.mask_dst, shape, True)
            self.reset_mask = load_float_mask(c.reset_mask, shape, 1)

        def dev(a, dtype):
            return torch.from_numpy(np.ascontiguousarray(a.astype(dtype))).cuda()
        m_alpha = None if c.mask_alpha is None else dev(np.asarray(self.mask_alpha, np.float32), np.float32)
        m_src = None if c.mask_src is None else dev(self.mask_src, np.uint8)
        m_dst = None if c.mask_dst is None else dev(self.mask_dst, np.uint8)
        scale = None
        if c.reset_mask is not None and c.reset_mode != "off":
            rm = np.asarray(self.reset_mask, np.float32)
            # evaluated by NumPy exactly as the reference does (python scalar * float32 array)
            if c.reset_mode == "random":
                scale = c.reset_random_factor * rm
            elif c.reset_mode == "constant":
                scale = c.reset_constant_step * rm
            else:
                scale = rm
            scale = dev(np.asarray(scale, np.float32), np.float32)
        check(self._lib.tf_layer_set_masks(self._handle, ptr(m_src), ptr(m_dst), ptr(m_alpha), ptr(scale),
                                           stream_ptr()))
        torch.cuda.current_stream().synchronize()  # the library copied the planes; temporaries may go

    def _push_sources(self):
        n = len(self.sources)
        masks = [torch.from_numpy(np.ascontiguousarray(np.asarray(s.introduction_mask, bool).astype(np.uint8))).cuda()
                 for s in self.sources]
        arr = (C.c_void_p * max(n, 1))(*[m.data_ptr() for m in masks])
        check(self._lib.tf_layer_set_sources(self._handle, n, arr, stream_ptr()))
        torch.cuda.current_stream().synchronize()

    def set_sources(self, sources):
        self.sources = list(sources)
        self._push_sources()

    # -- per frame ------------------------------------------------------------------------------
    def _pull_pixmaps(self):
        """``source.next()`` for every source, in order (reference.py:98, static.py:15)."""
        pixmaps = []
        for s in self.sources:
            pixmaps.append(_to_device_u8(s.next()))
        return pixmaps

    def _needs_pixmaps(self) -> bool:
        return True

    def takes_claims(self) -> bool:
        """Can ``Compositor.step`` feed this layer a forward flow as ``ops.ForwardClaims`` (the scatter pass's claim
        plane) instead of the flow?  Single-source move-reference layers on the fast kernel, device-side reset draws."""
        return (self.KIND == "moveref" and self.reset_rng != "numpy" and
                bool(self._lib.tf_layer_takes_claims(self._handle, len(self.sources))))

    def _update_claims(self, claims, rgb_inout, background=0):
        pixmaps = self._pull_pixmaps()
        arr = (PixmapStruct * 1)()
        pm = pixmaps[0]
        if tuple(pm.shape[:2]) != (self.height, self.width):
            raise ValueError(f"pixmap must be ({self.height}, {self.width}, C), got {tuple(pm.shape)}")
        arr[0].pixels = pm.data_ptr()
        arr[0].channels = int(pm.shape[2])
        arr[0].frame_number = int(self.sources[0].frame_number)
        plane = claims.take()
        check(self._lib.tf_layer_update_claims(self._handle, ptr(plane), arr, 1, self.rng_seed, ptr(rgb_inout),
                                               int(background), stream_ptr()))
        self._keepalive = (plane, pixmaps, None)

    def _update(self, flow, rgb_inout=None, first_layer=False, background=0):
        if isinstance(flow, ForwardClaims):       # outside Compositor.step's claim path: the flow itself
            flow = flow.tensor()
        fl = None if self.KIND == "static" else flow_to_device(flow)
        if fl is not None and tuple(fl.shape) != (self.height, self.width, 2):
            raise ValueError(f"flow must be ({self.height}, {self.width}, 2), got {tuple(fl.shape)}")
        pixmaps = self._pull_pixmaps() if self._needs_pixmaps() else []
        n = len(pixmaps)
        arr = (PixmapStruct * max(n, 1))()
        for i, pm in enumerate(pixmaps):
            if tuple(pm.shape[:2]) != (self.height, self.width):
                raise ValueError(f"pixmap must be ({self.height}, {self.width}, C), got {tuple(pm.shape)}")
            arr[i].pixels = pm.data_ptr()
            arr[i].channels = int(pm.shape[2])
            arr[i].frame_number = int(self.sources[i].frame_number)
        random = None
        if self.config.reset_mode == "random" and self.KIND in ("moveref", "sum") and self.reset_rng == "numpy":
            random = torch.from_numpy(np.random.random(size=(self.height, self.width))).cuda(non_blocking=True)
        n_expected = len(self.sources) if self._needs_pixmaps() else 0
        if n_expected == 0 and self.sources:
            # introduce_once already satisfied: the library expects the source count to match
            arr = (PixmapStruct * len(self.sources))()
            for i in range(len(self.sources)):
                arr[i].pixels = self._dummy_pixmap().data_ptr()
                arr[i].channels = 3
                arr[i].frame_number = int(self.sources[i].frame_number)
            n = len(self.sources)
        check(self._lib.tf_layer_update(self._handle, ptr(fl), arr, n, ptr(random), self.rng_seed, ptr(rgb_inout),
                                        int(bool(first_layer)), int(background), stream_ptr()))
        # keep inputs alive until the stream has consumed them
        self._keepalive = (fl, pixmaps, random)

    def _dummy_pixmap(self):
        if getattr(self, "_dummy", None) is None:
            self._dummy = torch.zeros((self.height, self.width, 3), dtype=torch.uint8, device="cuda")
        return self._dummy

    def update(self, flow):
        self._update(flow)

    def render(self) -> np.ndarray:
        return self.render_device().cpu().numpy()

    def render_device(self) -> torch.Tensor:
        out = torch.empty((self.height, self.width, 4), dtype=torch.uint8, device="cuda")
        check(self._lib.tf_layer_render(self._handle, ptr(out), stream_ptr()))
        return out

    def check_indices(self):
        """Raise IndexError if any flow vector pointed outside the frame (NumPy would have)."""
        check(self._lib.tf_layer_poll_error(self._handle, stream_ptr()))

    # -- state as the reference's arrays -------------------------------------------------------------
    @property
    def base(self) -> np.ndarray:
        return np.indices((self.height, self.width), dtype=np.int32).transpose(1, 2, 0)

    def _get_state(self):
        depth = self._lib.tf_layer_depth(self._handle)
        data = None if self.KIND == "static" else torch.empty((self.height, self.width, depth), dtype=torch.int32,
                                                                device="cuda")
        rgba = None if self.KIND == "introduction" else torch.empty((self.height, self.width, 4), dtype=torch.uint8,
                                                                      device="cuda")
        check(self._lib.tf_layer_get_state(self._handle, ptr(data), ptr(rgba), stream_ptr()))
        return (None if data is None else data.cpu().numpy(), None if rgba is None else rgba.cpu().numpy())

    @property
    def data(self):
        return self._get_state()[0]

    @data.setter
    def data(self, value):
        t = torch.from_numpy(np.ascontiguousarray(value, dtype=np.int32)).cuda()
        check(self._lib.tf_layer_set_state(self._handle, ptr(t), ptr(None), stream_ptr()))
        torch.cuda.current_stream().synchronize()

    @property
    def rgba(self):
        data, rgba = self._get_state()
        return data[:, :, :4] if self.KIND == "introduction" else rgba

    @rgba.setter
    def rgba(self, value):
        t = torch.from_numpy(np.ascontiguousarray(value, dtype=np.uint8)).cuda()
        check(self._lib.tf_layer_set_state(self._handle, ptr(None), ptr(t), stream_ptr()))
        torch.cuda.current_stream().synchronize()

    def __getstate__(self):
        data, rgba = self._get_state()
        frames, once = C.c_uint64(), C.c_int()
        check(self._lib.tf_layer_get_counters(self._handle, C.byref(frames), C.byref(once)))
        return {"config": self.config, "height": self.height, "width": self.width, "sources": self.sources,
                "data": data, "rgba": rgba, "frames": frames.value, "introduced_once": bool(once.value),
                "reset_rng": self.reset_rng, "rng_seed": self.rng_seed,
                "masks": (self.mask_alpha, self.mask_src, self.mask_dst, self.reset_mask)}

    def __setstate__(self, state):
        self.config = state["config"]
        self.height, self.width = state["height"], state["width"]
        self.sources = state["sources"]
        self.reset_rng, self.rng_seed = state["reset_rng"], state["rng_seed"]
        self._lib = _lib.load()
        self._handle = C.c_void_p()
        self._create()
        self._upload_masks(state.get("masks"))
        if self.sources:
            self._push_sources()
        d = None if state["data"] is None else torch.from_numpy(state["data"]).cuda()
        r = None if state["rgba"] is None else torch.from_numpy(state["rgba"]).cuda()
        check(self._lib.tf_layer_set_state(self._handle, ptr(d), ptr(r), stream_ptr()))
        check(self._lib.tf_layer_set_counters(self._handle, state["frames"], int(state["introduced_once"])))
        torch.cuda.current_stream().synchronize()

    def close(self):
        if getattr(self, "_handle", None) is not None and self._handle.value:
            self._lib.tf_layer_destroy(self._handle)
            self._handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @classmethod
    def from_args(cls, config: LayerConfig, height: int, width: int, sources):
        args = (config, height, width, sources)
        if config.classname == "moveref":
            from .move_reference import MoveReferenceLayer
            return MoveReferenceLayer(*args)
        if config.classname == "introduction":
            from .introduction import IntroductionLayer
            return IntroductionLayer(*args)
        if config.classname == "static":
            from .static import StaticLayer
            return StaticLayer(*args)
        if config.classname == "sum":
            from .sum import SumLayer
            return SumLayer(*args)
        if str(config.classname).split(":")[0] == "floatmap":     # extension
            from .floatmap import FloatMapLayer
            return FloatMapLayer(*args)
        raise ValueError(f"Unknown layer classname {config.classname}")
