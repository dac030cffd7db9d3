"""NumPy restatement of the reference compositor (``transflow/compositor/**``).

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).  The reference expresses every layer
with ``numpy.nonzero`` + ``.flat`` fancy indexing; this file restates the same semantics as
explicit per-pixel predicates (the form the CUDA kernels use), so that agreement with the
golden vectors produced by the real reference (``tests/golden/make_golden.py``) proves the
per-pixel reading of the reference is right, quirks included.

Record layouts (``compositor/layers/data.py:8-12``, ``introduction.py:10-14``):
  reference layers : int32 (H, W, 4) = (i, j, alpha, source)
  introduction     : int32 (H, W, 8) = (r, g, b, alpha, source, i, j, frame)
"""
from dataclasses import dataclass, field

import numpy as np


@dataclass
class LayerSpec:
    """Subset of ``LayerConfig`` (``transflow/config.py:57-104``) that reaches the arithmetic."""
    classname: str = "moveref"
    transparent_pixels_can_move: bool = False
    pixels_can_move_to_empty_spot: bool = True
    pixels_can_move_to_filled_spot: bool = True
    moving_pixels_leave_empty_spot: bool = False
    reset_mode: str = "off"
    reset_random_factor: float = 1
    reset_constant_step: float = 1
    reset_linear_factor: float = 0.1
    reset_source: bool = False
    introduce_pixels_on_empty_spots: bool = True
    introduce_pixels_on_filled_spots: bool = True
    introduce_moving_pixels: bool = True
    introduce_unmoving_pixels: bool = True
    introduce_once: bool = False
    introduce_on_all_filled_spots: bool = False
    introduce_on_all_empty_spots: bool = False


def flat_offsets(flow: np.ndarray, width: int) -> np.ndarray:
    """``MovementLayer._update_flow`` (movement.py:20-23): half-even round -> int32 -> fy*W+fx."""
    fi = np.rint(flow).astype(np.int32)
    return (fi[..., 1] * np.int32(width) + fi[..., 0]).ravel()


def move_records(data, flow, mask_src, mask_dst, spec: LayerSpec, alpha_index: int):
    """``MovementLayer._update_move`` (movement.py:25-60) as one gather.

    Returns the new ``data`` (a fresh array).  Source index q = p + off uses NumPy ``.flat``
    semantics (negative indices wrap; q >= H*W raises IndexError, as in the reference).
    """
    h, w, depth = data.shape
    n = h * w
    off = flat_offsets(flow, w)
    q = np.arange(n) + off
    alpha_old = data[..., alpha_index].ravel()
    src_alpha = alpha_old.flat[q] != 0          # raises like the reference when q >= n
    src_ok = mask_src.ravel().flat[q].astype(bool)
    if not spec.transparent_pixels_can_move:
        src_ok &= src_alpha
    dst_ok = mask_dst.ravel().astype(bool).copy()
    if not spec.pixels_can_move_to_empty_spot:
        dst_ok &= alpha_old != 0
    if not spec.pixels_can_move_to_filled_spot:
        dst_ok &= alpha_old == 0
    target = (off != 0) & src_ok & dst_ok
    old = data.reshape(n, depth)
    new = old.copy()
    new[target] = old[q[target]]
    alpha_new = new[:, alpha_index]
    if spec.moving_pixels_leave_empty_spot:
        alpha_new[q[target]] = 0
    if spec.transparent_pixels_can_move:
        alpha_new[target & src_alpha] = 1
    else:
        alpha_new[target] = 1
    return new.reshape(h, w, depth)


def base_source_plane(intro_masks) -> np.ndarray:
    """Last source whose introduction mask covers the pixel, -1 if none
    (``ReferenceLayer._set_base_source_indices``, reference.py:46-52 applies them in order)."""
    if not intro_masks:
        return None
    out = np.full(intro_masks[0].shape, -1, np.int32)
    for s, m in enumerate(intro_masks):
        out[np.asarray(m, bool)] = s
    return out


def reset_random(data, base, threshold_f32, random_f64, base_src, reset_source: bool):
    """``_update_reset_random`` (reference.py:58-67).  ``threshold_f32`` is
    ``reset_random_factor * reset_mask`` evaluated by NumPy (float32); the compare is float64."""
    hit = random_f64 < threshold_f32
    data[..., 0][hit] = base[..., 0][hit]
    data[..., 1][hit] = base[..., 1][hit]
    data[..., 2][hit] = 1
    if reset_source and base_src is not None:
        sel = hit & (base_src >= 0)
        data[..., 3][sel] = base_src[sel]
    return data


def reset_constant(data, base, step, reset_mask):
    """``_update_reset_constant`` (reference.py:69-79), float32 throughout."""
    d0 = (base - data[..., :2]).astype(np.float32)
    n0 = np.max(np.abs(d0), axis=2)
    d = d0.copy()
    nz = n0 != 0
    d[nz] = d[nz] / n0[nz][:, None]
    d = d * (step * reset_mask)[..., None]          # python scalar * float32 -> float32
    n1 = np.max(np.abs(d), axis=2)
    over = n1 > n0
    d[over] = d0[over]
    data[..., :2] += np.rint(d).astype(np.int32)
    return data


def reset_linear(data, base, factor, reset_mask):
    """``_update_reset_linear`` (reference.py:81-83): float64 product, half-even round."""
    d = factor * (base - data[..., :2])              # float64
    data[..., :2] += np.rint(reset_mask[..., None] * d).astype(np.int32)
    return data


def remap_rgba(data, rgba, pixmaps):
    """``ReferenceLayer._update_rgba`` (reference.py:93-105), per source in order."""
    h, w = data.shape[:2]
    for s, pm in enumerate(pixmaps):
        sel = (data[..., 3] == s) & (data[..., 2] != 0)
        ii = np.clip(data[..., 0], 0, h - 1)
        jj = np.clip(data[..., 1], 0, w - 1)
        c = pm.shape[2]
        rgba[..., :c][sel] = pm[ii[sel], jj[sel]]
        if c == 3:
            rgba[..., 3] = sel.astype(np.uint8)
    return rgba


class LayerOracle:
    """One compositor layer.  ``update(flow, pixmaps, random=None)``, ``render()``."""

    def __init__(self, spec: LayerSpec, height: int, width: int, intro_masks=(),
                 mask_src=None, mask_dst=None, mask_alpha=None, reset_mask=None):
        self.spec, self.h, self.w = spec, height, width
        self.kind = spec.classname
        self.intro_masks = [np.asarray(m, bool) for m in intro_masks]
        ones_b = np.ones((height, width), bool)
        self.mask_src = ones_b if mask_src is None else np.asarray(mask_src, bool)
        self.mask_dst = ones_b if mask_dst is None else np.asarray(mask_dst, bool)
        self.mask_alpha = (np.ones((height, width), np.float32) if mask_alpha is None
                           else np.asarray(mask_alpha, np.float32))
        self.reset_mask = (np.ones((height, width), np.float32) if reset_mask is None
                           else np.asarray(reset_mask, np.float32))
        self.base = np.indices((height, width), dtype=np.int32).transpose(1, 2, 0)
        self.rgba = np.zeros((height, width, 4), np.uint8)
        self.frame_numbers = [-1] * len(self.intro_masks)
        self.introduced_once = False
        if self.kind in ("moveref", "sum"):
            self.data = np.zeros((height, width, 4), np.int32)
            self.data[..., :2] = self.base
            self.data[..., 2] = 1
            self.base_src = base_source_plane(self.intro_masks)
            if self.base_src is not None:
                sel = self.base_src >= 0
                self.data[..., 3][sel] = self.base_src[sel]
        elif self.kind == "introduction":
            self.data = np.zeros((height, width, 8), np.int32)
        elif self.kind == "static":
            self.data = None
            self.rgba[..., 3] = 1
        else:
            raise ValueError(f"Unknown layer classname {self.kind}")

    # -- reference layers ------------------------------------------------------------
    def _reset(self, random):
        sp = self.spec
        if sp.reset_mode == "random":
            if random is None:
                random = np.random.random(size=(self.h, self.w))
            thr = sp.reset_random_factor * self.reset_mask
            reset_random(self.data, self.base, thr, random, self.base_src, sp.reset_source)
        elif sp.reset_mode == "constant":
            reset_constant(self.data, self.base, sp.reset_constant_step, self.reset_mask)
        elif sp.reset_mode == "linear":
            reset_linear(self.data, self.base, sp.reset_linear_factor, self.reset_mask)
        elif sp.reset_mode != "off":
            raise ValueError(f"Unknown reset mode {sp.reset_mode}")

    # -- introduction (introduction.py:20-67) ----------------------------------------
    def _introduce(self, flow, pixmaps):
        sp = self.spec
        if sp.introduce_once and self.introduced_once:
            return
        self.introduced_once = True
        n = self.h * self.w
        off = flat_offsets(flow, self.w)
        alpha = self.data[..., 3].ravel()
        ok = np.ones(n, bool)
        # Q13: the "empty spot" and "unmoving" switches index with ``numpy.where(x) == 0``
        # (a tuple compared to 0 -> False) and therefore never mask anything.
        if not sp.introduce_pixels_on_filled_spots:
            ok &= alpha == 0
        if not sp.introduce_moving_pixels:
            ok &= off == 0
        if sp.introduce_on_all_filled_spots:
            ok |= alpha != 0
        use_flow = not (sp.introduce_on_all_filled_spots or sp.introduce_on_all_empty_spots)
        flat = self.data.reshape(n, 8)
        for s, pm in enumerate(pixmaps):
            self.frame_numbers[s] += 1               # PixmapSourceInterface.next() bumps counter
            tgt = np.nonzero(ok & self.intro_masks[s].ravel())[0]
            src = tgt + off[tgt] if use_flow else tgt
            pix = pm.reshape(n, pm.shape[2])
            rec = np.empty((tgt.size, 8), np.int32)
            rec[:, 0:3] = pix[:, :3][src]            # negative src wraps, like .flat[]
            rec[:, 3] = pix[:, 3][src] if pm.shape[2] == 4 else 1
            rec[:, 4] = s
            rec[:, 5:7] = self.base.reshape(n, 2)[src]
            rec[:, 7] = self.frame_numbers[s]
            flat[tgt] = rec

    def update(self, flow, pixmaps=(), random=None):
        flow = np.asarray(flow, np.float32)
        pixmaps = list(pixmaps)
        if self.kind == "moveref":
            self.data = move_records(self.data, flow, self.mask_src, self.mask_dst, self.spec, 2)
            self._reset(random)
            remap_rgba(self.data, self.rgba, pixmaps)
        elif self.kind == "sum":
            # Q8: flow x is added to the ROW index, flow y to the COLUMN index (sum.py:10)
            self.data[..., :2] += np.floor(flow).astype(np.int32)
            self._reset(random)
            remap_rgba(self.data, self.rgba, pixmaps)
        elif self.kind == "static":
            for s, pm in enumerate(pixmaps):
                m = self.intro_masks[s]
                self.rgba[..., :pm.shape[2]][m] = pm[m]
        elif self.kind == "introduction":
            self.data = move_records(self.data, flow, self.mask_src, self.mask_dst, self.spec, 3)
            self._introduce(flow, pixmaps)

    def render(self) -> np.ndarray:
        """``Layer.render`` (layer.py:32-34): alpha *= mask_alpha IN PLACE (truncating), clip, u8."""
        if self.kind == "introduction":
            a = self.data[..., 3]
            a[...] = (self.mask_alpha * a).astype(np.int32)
            return np.clip(self.data[..., :4], 0, 255).astype(np.uint8)
        a = self.rgba[..., 3]
        a[...] = (self.mask_alpha * a).astype(np.uint8)
        return self.rgba.copy()


def composite(background_rgb: np.ndarray, layer_images) -> np.ndarray:
    """``Compositor.render`` (compositor.py:31-40): opaque overwrite in layer order."""
    out = background_rgb.copy()
    for img in layer_images:
        opaque = img[..., 3] != 0
        out[opaque] = img[..., :3][opaque]
    return out


class CompositorOracle:
    def __init__(self, height, width, layers, background_rgb=(255, 255, 255)):
        self.layers = list(layers)
        self.background = np.empty((height, width, 3), np.uint8)
        self.background[:, :] = background_rgb

    def update(self, flow, pixmaps_per_layer=None, randoms=None):
        for li, layer in enumerate(self.layers):
            pm = () if pixmaps_per_layer is None else pixmaps_per_layer.get(li, ())
            rnd = None if randoms is None else randoms.get(li)
            layer.update(flow, pm, rnd)

    def render(self):
        return composite(self.background, [layer.render() for layer in self.layers])
