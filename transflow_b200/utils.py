"""Host-side setup helpers shared by the flow and compositor plugins: the mask DSL, colour
parsing, lambda expressions.  These run once per pipeline (not per frame) and stay NumPy, so
masks are bit-identical to the reference's (``transflow/utils.py:51-144``, ``:316-324``).
"""
import re
import warnings
from typing import Callable

import numpy  # noqa: F401  (user expressions may name it, as in transflow/utils.py)
import numpy as np


def parse_dimension_arg(arg: str, parent_size: int) -> int:
    """``"12"`` -> 12 px, ``"25%"`` -> int(0.25 * parent), empty -> 0."""
    arg = arg.strip()
    if not arg:
        return 0
    if arg.endswith("%"):
        return int(float(arg[:-1]) / 100 * parent_size)
    return int(arg)


def _border_sizes(rule: str, height: int, width: int):
    """-> (top, right, bottom, left) for ``border[:-side]:a[:b[:c:d]]`` (CSS order)."""
    name, args = rule.lower().split(":", 1)
    sides = {"border-top": 0, "border-right": 1, "border-bottom": 2, "border-left": 3}
    out = [0, 0, 0, 0]
    if name == "border":
        vals = [parse_dimension_arg(a, height if i % 2 == 0 else width) for i, a in enumerate(args.split(":"))]
        if len(vals) == 1:
            out = [vals[0]] * 4
        elif len(vals) == 2:
            out = [vals[0], vals[1], vals[0], vals[1]]
        elif len(vals) == 4:
            out = vals
        else:
            raise ValueError(f"Invalid number of argument {len(vals)} for border mask")
    elif name in sides:
        i = sides[name]
        out[i] = parse_dimension_arg(args, height if i % 2 == 0 else width)
    else:
        raise ValueError(f"Invalid border rule name {name}")
    return out


_BORDER_RE = re.compile(r"^border(\-(top|right|bottom|left))?:(\d+%?:|:|\d+%?$){1,4}$", re.IGNORECASE)


def load_float_mask(mask_path: str | None, shape: tuple[int, int] = (0, 0), default: float = 0) -> np.ndarray:
    """Mask DSL of the reference (utils.py:51-140) -> float32 (H, W) in [0, 1].

    ``None`` -> constant ``default``; ``zeros`` / ``ones`` / ``random``; ``border...``,
    ``hline:N`` / ``vline:N``, ``circle:R``, ``rect:W[:H]``, ``grid:rows:cols:radius``; anything
    else is an image path (grey or mean of RGB, /255).  A trailing ``:inv`` inverts.
    """
    if mask_path is None:
        return np.full(shape, default, dtype=np.float32)
    spec = mask_path
    inverse = spec.endswith(":inv")
    if inverse:
        spec = spec[:-4]
    low = spec.lower()
    h, w = shape
    if low == "zeros":
        arr = np.zeros(shape, np.float32)
    elif low == "ones":
        arr = np.ones(shape, np.float32)
    elif low == "random":
        arr = np.random.rand(*shape).astype(np.float32)
    elif _BORDER_RE.match(spec):
        top, right, bottom, left = _border_sizes(spec, h, w)
        arr = np.zeros(shape, np.float32)
        if top:
            arr[:top, :] = 1
        if right:
            arr[:, -right:] = 1
        if bottom:
            arr[-bottom:, :] = 1
        if left:
            arr[:, :left] = 1
    elif re.match(r"^[hv]line:\d+%?$", spec, re.IGNORECASE):
        name, arg = low.split(":")
        arr = np.zeros(shape, np.float32)
        if name == "hline":
            n = parse_dimension_arg(arg, h)
            i = (h - n) // 2
            arr[i:i + n, :] = 1
        else:
            n = parse_dimension_arg(arg, w)
            j = (w - n) // 2
            arr[:, j:j + n] = 1
    elif re.match(r"circle:\d+%?", spec, re.IGNORECASE):
        radius = parse_dimension_arg(low.split(":")[1], min(shape))
        ii = np.arange(h)[:, None] - h // 2
        jj = np.arange(w)[None, :] - w // 2
        arr = jj ** 2 + ii ** 2 < radius ** 2          # bool, like the reference
    elif re.match(r"rect:\d+%?(:\d+%?)?", spec, re.IGNORECASE):
        args = spec[spec.index(":") + 1:].split(":")
        if len(args) == 1:
            rw, rh = parse_dimension_arg(args[0], w), parse_dimension_arg(args[0], h)
        elif len(args) == 2:
            rw, rh = parse_dimension_arg(args[0], w), parse_dimension_arg(args[1], h)
        else:
            raise ValueError(f"Invalid number of argument {len(args)} for rect mask")
        arr = np.ones(shape, np.float32)
        arr[:h // 2 - rh // 2, :] = 0
        arr[h // 2 + rh // 2:, :] = 0
        arr[:, :w // 2 - rw // 2] = 0
        arr[:, w // 2 + rw // 2:] = 0
    elif re.match(r"grid:\d+:\d+:\d+?", spec, re.IGNORECASE):
        nrows, ncols, radius = (int(a) for a in spec[spec.index(":") + 1:].split(":"))
        d = 2 * radius
        k = np.arange(d) - radius
        disc = k[None, :] ** 2 + k[:, None] ** 2 < radius ** 2
        arr = np.zeros(shape, np.float32)
        ch, cw = h // nrows, w // ncols
        for r in range(nrows):
            for c in range(ncols):
                i0 = ch * r + ch // 2 - radius
                j0 = cw * c + cw // 2 - radius
                arr[i0:i0 + d, j0:j0 + d] = disc
    else:
        import PIL.Image
        with PIL.Image.open(spec) as image:
            arr = np.array(image).astype(np.float32)
        if arr.ndim == 2:
            arr /= 255
        elif arr.ndim == 3:
            if arr.shape[2] == 4:
                warnings.warn(f"Mask {spec} has an alpha channel but it will be ignored")
            arr = np.mean(arr[:, :, :3], axis=2) / 255
        else:
            raise ValueError(f"Image has wrong number of dimensions {arr.ndim}, expected 2 or 3")
    if inverse:
        arr = 1.0 - arr
    return arr


def load_bool_mask(mask_path: str | None, shape: tuple[int, int] = (0, 0), default: bool = False) -> np.ndarray:
    """Rounded (half-even) float mask as bool (utils.py:143-144)."""
    return np.round(load_float_mask(mask_path, shape, float(default))).astype(bool)


def parse_color(string: str) -> tuple[int, int, int]:
    """CSS colour name, ``rgb(r, g, b)`` / ``(r, g, b)``, or hex (``#rrggbb``, ``0x...``)."""
    import PIL.ImageColor
    name = string.lower()
    if name in PIL.ImageColor.colormap:
        return tuple(PIL.ImageColor.getrgb(name)[:3])
    m = re.match(r"^(?:rgb)?\((\d+), ?(\d+), ?(\d+)\)$", string, re.IGNORECASE)
    if m:
        return (int(m.group(1)), int(m.group(2)), int(m.group(3)))
    x = int(string.replace("#", "").replace("0x", "").replace("x", ""), 16)
    return ((x >> 16) & 255, (x >> 8) & 255, x & 255)


def parse_lambda_expression(expr: str, variables: tuple[str, ...] = ("t",), numpy_module=None) -> Callable:
    """``"1 + t"`` -> ``lambda t: 1 + t`` (trusted CLI input, as in the reference).  The expression sees what the
    reference's ``utils`` module offers it (utils.py:1-8, USAGE.md: ``math``, ``random``, ``numpy``, ...)."""
    import math
    import os
    import random
    npm = np if numpy_module is None else numpy_module      # polar filters pass a tensor-aware stand-in
    scope = {"math": math, "random": random, "numpy": npm, "np": npm, "os": os, "re": re}
    return eval("lambda " + ",".join(variables) + ": " + expr, scope)  # noqa: S307


def parse_timestamp(value) -> float | None:
    """``None`` | number | ``"[[hh:]mm:]ss[.mmm]"`` -> seconds."""
    if value is None:
        return None
    if isinstance(value, (int, float)):
        return float(value)
    text = str(value).strip()
    try:
        return float(text)
    except ValueError:
        pass
    parts = text.split(":")
    seconds = 0.0
    for part in parts:
        seconds = seconds * 60 + float(part)
    return seconds


def upscale_array(arr: np.ndarray, wf: int, hf: int) -> np.ndarray:
    """Integer block upscale of a flow with vector scaling (utils.py:417-418)."""
    scaled = arr * np.asarray((wf, hf), dtype=arr.dtype)
    return np.repeat(np.repeat(scaled, hf, axis=0), wf, axis=1).astype(arr.dtype)
