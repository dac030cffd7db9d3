"""Oracle for the flow side of the hot path.  TEST INFRASTRUCTURE (see ``oracle/__init__.py``).

Farneback and pyramidal Lucas-Kanade are third-party arithmetic (``opencv-python``; this
image: opencv-python-headless 4.13.0.92).  The oracle calls ``cv2`` with exactly the
arguments of the reference's call sites; Horn-Schunck and the post-process step are
reference-owned code and are restated in NumPy.
"""
import numpy as np


def gray_from_bgr(bgr: np.ndarray) -> np.ndarray:
    """``cv2.cvtColor(BGR2GRAY)`` as called at ``flow/sources/cv.py:465``: 15-bit fixed point
    ``(3735 B + 19235 G + 9798 R + 16384) >> 15`` (bit-exact vs cv2 4.13)."""
    b = bgr[..., 0].astype(np.int32)
    g = bgr[..., 1].astype(np.int32)
    r = bgr[..., 2].astype(np.int32)
    return ((3735 * b + 19235 * g + 9798 * r + 16384) >> 15).astype(np.uint8)


def farneback(left: np.ndarray, right: np.ndarray, pyr_scale=0.5, levels=3, winsize=15,
              iterations=3, poly_n=5, poly_sigma=1.2, flags=0) -> np.ndarray:
    """Call site ``flow/sources/cv.py:477-490`` (flags=0: the passed flow buffer is ignored)."""
    import cv2
    flow = np.zeros(left.shape + (2,), np.float32)
    return cv2.calcOpticalFlowFarneback(prev=left, next=right, flow=flow, pyr_scale=pyr_scale,
                                        levels=levels, winsize=winsize, iterations=iterations,
                                        poly_n=poly_n, poly_sigma=poly_sigma,
                                        flags=flags).astype(np.float32)


def lucas_kanade(left: np.ndarray, right: np.ndarray, win_size=15, max_level=2, step=1) -> np.ndarray:
    """``flow/methods/lukas_kanade.py:9-36``: PyrLK on the grid of every ``step``-th pixel,
    status ignored, block-replicated back to (H, W)."""
    import cv2
    h, w = left.shape
    gx, gy = np.meshgrid(np.arange(0, w, step), np.arange(0, h, step), indexing="xy")
    p0 = np.stack([gx, gy], axis=-1).astype(np.float32)
    rows, cols = p0.shape[:2]
    p0 = p0.reshape(rows * cols, 1, 2)
    p1 = p0.copy()
    cv2.calcOpticalFlowPyrLK(left, right, p0, p1, winSize=(win_size, win_size), maxLevel=max_level)
    flow = (p1 - p0).reshape(rows, cols, 2)
    if step == 1:
        return flow
    return np.repeat(np.repeat(flow, step, axis=0), step, axis=1)[:h, :w].astype(np.float32)


def _blur5(img_f32: np.ndarray) -> np.ndarray:
    """``cv2.GaussianBlur(x, (5, 5), 0)``: separable [1 4 6 4 1]/16, BORDER_REFLECT_101."""
    k = np.asarray([1, 4, 6, 4, 1], np.float32) / np.float32(16)
    p = np.pad(img_f32, ((0, 0), (2, 2)), mode="reflect")
    t = sum(k[j] * p[:, j:j + img_f32.shape[1]] for j in range(5))
    p = np.pad(t, ((2, 2), (0, 0)), mode="reflect")
    return sum(k[j] * p[j:j + img_f32.shape[0], :] for j in range(5)).astype(np.float32)


def _fwd(x, axis):
    """x shifted by +1 along axis with the last sample repeated (scipy 'reflect' == symmetric)."""
    return np.concatenate([np.take(x, range(1, x.shape[axis]), axis=axis),
                           np.take(x, [x.shape[axis] - 1], axis=axis)], axis=axis)


def _avg3(u):
    """``scipy.ndimage.convolve(u, [[1,2,1],[2,0,2],[1,2,1]]/12)`` with symmetric borders."""
    p = np.pad(u, 1, mode="symmetric")
    h, w = u.shape
    s = (p[0:h, 0:w] + p[0:h, 2:w + 2] + p[2:h + 2, 0:w] + p[2:h + 2, 2:w + 2]
         + 2 * (p[0:h, 1:w + 1] + p[2:h + 2, 1:w + 1] + p[1:h + 1, 0:w] + p[1:h + 1, 2:w + 2]))
    return s / 12


def horn_schunck(left, right, flow=None, alpha=1, max_iters=3, decay=0, delta=1,
                 use_blur_from_cv2=True, sweeps_out=None):
    """``flow/methods/horn_schunck.py:9-45``.

    ``scipy.ndimage.convolve`` with a 2x2 kernel and mode='reflect' evaluates
    out[i,j] = sum_{p,q in {0,1}} k[p,q] * x[i+1-p, j+1-q] with symmetric borders
    (SURVEY.md A.3).  The early exit uses the matrix 2-norm (largest singular value) of
    the change in ``u`` only.  ``sweeps_out`` (list) receives the number of sweeps executed.
    """
    if use_blur_from_cv2:
        import cv2
        a = cv2.GaussianBlur(left.astype(np.float32), (5, 5), 0)
        b = cv2.GaussianBlur(right.astype(np.float32), (5, 5), 0)
    else:
        a, b = _blur5(left.astype(np.float32)), _blur5(right.astype(np.float32))
    if flow is None:
        u = np.zeros(a.shape)
        v = np.zeros(a.shape)
    else:
        u = decay * flow[..., 0]
        v = decay * flow[..., 1]

    def d2x2(x, kx):
        # kx[p][q] applies to x[i+1-p, j+1-q]
        x01, x10, x11 = _fwd(x, 1), _fwd(x, 0), _fwd(_fwd(x, 0), 1)
        return kx[0][0] * x11 + kx[0][1] * x10 + kx[1][0] * x01 + kx[1][1] * x
    q = 0.25
    xk = [[q, -q], [q, -q]]
    yk = [[q, q], [-q, -q]]
    tk = [[q, q], [q, q]]
    ex = (d2x2(a, xk) + d2x2(b, xk)).astype(np.float32)
    ey = (d2x2(a, yk) + d2x2(b, yk)).astype(np.float32)
    et = (d2x2(b, tk) - d2x2(a, tk)).astype(np.float32)
    n = 0
    for _ in range(max_iters):
        ua = _avg3(u).astype(u.dtype)
        va = _avg3(v).astype(v.dtype)
        c = (ex * ua + ey * va + et) / (alpha ** 2 + ex ** 2 + ey ** 2)
        prev = u
        u = ua - ex * c
        v = va - ey * c
        n += 1
        if delta is not None and np.linalg.norm(u - prev, 2) < delta:
            break
    if sweeps_out is not None:
        sweeps_out.append(n)
    return np.stack([u, v], axis=-1).astype(np.float32)


def apply_filter(flow: np.ndarray, name: str, value, t: float = 0.0) -> None:
    """One flow filter, in place (``flow/filters.py:36-87``).  ``value`` is the per-frame scalar the
    reference gets from ``expr(t)`` (for ``polar``: a pair of callables of ``(t, r, a)``)."""
    h, w, _ = flow.shape
    if name == "scale":                                   # filters.py:41
        flow *= value
        return
    norm = np.linalg.norm(flow.reshape(h * w, 2), axis=1).reshape((h, w))
    if name == "threshold":                               # filters.py:49-53
        flow[np.where(norm <= value)] = 0
    elif name == "clip":                                  # filters.py:61-68
        factors = np.ones((h, w))
        where = np.where(norm >= value)
        factors[where] = value / norm[where]
        flow[:, :, 0] *= factors
        flow[:, :, 1] *= factors
    elif name == "polar":                                 # filters.py:80-86
        theta = np.atan2(flow[:, :, 1], flow[:, :, 0])
        new_radius = value[0](t, norm, theta)
        new_theta = value[1](t, norm, theta)
        flow[:, :, 1] = new_radius * np.sin(new_theta)
        flow[:, :, 0] = new_radius * np.cos(new_theta)
    else:
        raise ValueError(f"Unknown filter name '{name}'")


def post_process(flow: np.ndarray, forward: bool, mask=None, kernel=None, filters=(), t: float = 0.0) -> np.ndarray:
    """``FlowSource.post_process`` (``flow/sources/source.py:337-363``).

    filters ([(name, scalar), ...], source.py:339-341) -> mask (:342-343) -> convolution kernel through the
    reference's own ``scipy.signal.convolve2d`` call (:344-348; the flow is float64 from there on) ->
    forward: clip, half-even round, scatter source coordinates to their targets (last writer
    in raster order wins), flow := origin - target; then the final clip (always).  Mutates
    ``flow`` in place like the reference unless a mask / kernel forces a copy.
    """
    for name, value in filters:
        apply_filter(flow, name, value, t)
    h, w = flow.shape[:2]
    if mask is not None:
        flow = np.multiply(np.asarray(mask, np.float32).reshape(h, w, 1), flow)
    if kernel is not None:
        import scipy.signal
        fx = scipy.signal.convolve2d(flow[:, :, 0], kernel, mode="same", boundary="fill", fillvalue=0)
        fy = scipy.signal.convolve2d(flow[:, :, 1], kernel, mode="same", boundary="fill", fillvalue=0)
        flow = np.stack([fx, fy], axis=-1)
    xs = np.arange(w, dtype=np.int32)[None, :]
    ys = np.arange(h, dtype=np.int32)[:, None]
    lo_x, hi_x, lo_y, hi_y = -xs, w - 1 - xs, -ys, h - 1 - ys
    if forward:
        np.clip(flow[..., 0], lo_x, hi_x, out=flow[..., 0])
        np.clip(flow[..., 1], lo_y, hi_y, out=flow[..., 1])
        fi = np.rint(flow).astype(np.int32)
        off = (fi[..., 1] * w + fi[..., 0]).ravel()
        src = np.nonzero(off)[0]
        owner = np.arange(h * w)
        owner[src + off[src]] = src          # duplicate targets: last (largest) source wins
        flow[..., 0] = (owner % w).reshape(h, w) - xs
        flow[..., 1] = (owner // w).reshape(h, w) - ys
    np.clip(flow[..., 0], lo_x, hi_x, out=flow[..., 0])
    np.clip(flow[..., 1], lo_y, hi_y, out=flow[..., 1])
    return flow


def forward_claims(flow: np.ndarray, mask=None, filters=(), t: float = 0.0) -> np.ndarray:
    """The claim plane behind the forward direction of ``post_process`` (``source.py:349-360``, the ``numpy.put`` of the
    source coordinates): int32 (H, W), ``claims[q] = p + 1`` for the LAST pixel p in raster order whose clipped, rounded
    vector points at q (p != q), 0 where nobody does.  ``post_process(flow, True)`` is exactly ``claimant - position``
    (0 where unclaimed), which is what the device's ``tf_flow_forward_claims`` / ``tf_layer_update_claims`` pair relies on
    (DESIGN.md 4a).  ``flow`` is not modified."""
    flow = np.array(flow, dtype=np.float32, copy=True)
    for name, value in filters:
        apply_filter(flow, name, value, t)
    h, w = flow.shape[:2]
    if mask is not None:
        flow = np.multiply(np.asarray(mask, np.float32).reshape(h, w, 1), flow)
    xs = np.arange(w, dtype=np.int32)[None, :]
    ys = np.arange(h, dtype=np.int32)[:, None]
    np.clip(flow[..., 0], -xs, w - 1 - xs, out=flow[..., 0])
    np.clip(flow[..., 1], -ys, h - 1 - ys, out=flow[..., 1])
    fi = np.rint(flow).astype(np.int32)
    off = (fi[..., 1] * w + fi[..., 0]).ravel()
    src = np.nonzero(off)[0]
    claims = np.zeros(h * w, dtype=np.int32)
    claims[src + off[src]] = src + 1         # duplicate targets: last (largest) source wins
    return claims.reshape(h, w)


def merge_flows(flows, mode: str) -> np.ndarray:
    """``Pipeline.FLOW_MERGING_FUNCTIONS[mode](flows)`` (``pipeline.py:149-158``, helpers ``utils.py:359-381``).
    Like the reference, ``maskbin`` overwrites the extra flows in place."""
    def multiply_arrays(arrays):                                  # utils.py:359-365
        if len(arrays) == 1:
            return arrays[0]
        out = np.multiply(arrays[0], arrays[1])
        for a in arrays[2:]:
            np.multiply(out, a, out)
        return out

    def binarize_arrays(arrays):                                  # utils.py:368-373
        for a in arrays:
            where = np.where(np.abs(a) > 0.2)
            a[:, :] = 0
            a[where] = 1
        return arrays

    def absmax(arrays):                                           # utils.py:376-381
        w, h = arrays[0].shape[0:2]
        stack = np.stack(arrays).reshape((2, w * h * 2))
        argmax = np.argmax(np.abs(stack), axis=0).reshape((1, w * h * 2))
        return np.take_along_axis(stack, argmax, 0).reshape(arrays[0].shape)

    table = {
        "first": lambda f: f[0],
        "sum": lambda f: np.sum(f, axis=0),
        "average": lambda f: np.sum(f, axis=0) / len(f),
        "difference": lambda f: f[0] - sum(f[1:]),
        "product": multiply_arrays,
        "maskbin": lambda f: multiply_arrays([f[0]] + binarize_arrays(f[1:])),
        "masklin": lambda f: multiply_arrays([f[0]] + [np.abs(x) for x in f[1:]]),
        "absmax": absmax,
    }
    return table[mode](list(flows))


def upscale_array(arr: np.ndarray, wf: int, hf: int) -> np.ndarray:
    """``utils.upscale_array`` (``utils.py:417-418``)."""
    return np.kron(arr * (wf, hf), np.ones((hf, wf, 1))).astype(arr.dtype)


# ------------------------------------------------------------------------------------------------
# flow visualisers (transflow/output/render.py:9-48; called at pipeline.py:509-516)
# ------------------------------------------------------------------------------------------------
def _parse_color(string: str):
    """``utils.parse_color`` (utils.py:316-324) for the forms the tests use (hex and rgb(...))."""
    import re
    m = re.match(r"^(?:rgb)?\((\d+), ?(\d+), ?(\d+)\)$", string, re.IGNORECASE)
    if m:
        return (int(m.group(1)), int(m.group(2)), int(m.group(3)))
    x = int(string.replace("#", "").replace("0x", "").replace("x", ""), 16)
    return ((x >> 16) & 255, (x >> 8) & 255, x & 255)


def render1d(arr: np.ndarray, scale=1, colors=None, binary=False) -> np.ndarray:
    """``render1d`` (render.py:9-28): two-colour ramp of a scalar array, float32 arithmetic."""
    if colors is None:
        colors = ("#000000", "#ffffff")
    c = [np.array(_parse_color(s), dtype=np.float32) for s in colors]
    shape = (*arr.shape[:2], 1)
    if binary:
        b = np.clip(np.round(scale * arr), 0, 1).reshape(shape)
        a = 1 - b
    else:
        a = np.clip(1 - scale * arr, 0, 1).reshape(shape)
        b = np.clip(scale * arr, 0, 1).reshape(shape)
    frame = np.multiply(a, c[0]) + np.multiply(b, c[1])
    return np.clip(frame, 0, 255).astype(np.uint8)


def render2d(arr: np.ndarray, scale=1, colors=None) -> np.ndarray:
    """``render2d`` (render.py:31-48): four-colour chart of a flow field."""
    if colors is None:
        colors = ("#ffff00", "#0000ff", "#ff00ff", "#00ff00")
    c = [np.array(_parse_color(s), dtype=np.float32) for s in colors]
    shape = (*arr.shape[:2], 1)
    cy = np.clip(1 + scale * arr[:, :, 0], 0, 1).reshape(shape)
    cb = np.clip(1 - scale * arr[:, :, 0], 0, 1).reshape(shape)
    cm = np.clip(1 + scale * arr[:, :, 1], 0, 1).reshape(shape)
    cg = np.clip(1 - scale * arr[:, :, 1], 0, 1).reshape(shape)
    frame = .5 * (np.multiply(cy, c[0]) + np.multiply(cb, c[1]) + np.multiply(cm, c[2]) + np.multiply(cg, c[3]))
    return np.clip(frame, 0, 255).astype(np.uint8)


def flow_magnitude(flow: np.ndarray) -> np.ndarray:
    """``numpy.sqrt(numpy.sum(numpy.power(flow, 2), axis=2))`` (pipeline.py:515)."""
    return np.sqrt(np.sum(np.power(flow, 2), axis=2))
