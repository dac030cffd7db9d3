"""CPU: pin the oracle against the golden vectors produced by the real reference and
against the live cv2 build."""
import os

import numpy as np
import pytest

from oracle import compositor_np as C
from oracle import farneback_np as FB
from tests import golden_util as G

Z = G.load("compositor_golden.npz")
MASK_KEYS = ("mask_alpha", "mask_src", "mask_dst", "reset_mask")
SPEC_KEYS = set(C.LayerSpec.__dataclass_fields__)


def build_oracle_layers(z, name):
    layers = []
    for li, kw in enumerate(G.case_layers(z, name)):
        spec = C.LayerSpec(**{k: v for k, v in kw.items() if k in SPEC_KEYS})
        masks = {k: z[f"{name}/{k}{li}"] for k in MASK_KEYS if f"{name}/{k}{li}" in z.files}
        srcs = G.case_sources(z, name, li)
        h, w = z[f"{name}/flows"].shape[1:3]
        layers.append(C.LayerOracle(spec, h, w, intro_masks=[m for _, m in srcs], **masks))
    return layers


@pytest.mark.parametrize("name", G.compositor_case_names(Z))
def test_compositor_oracle_matches_reference(name):
    flows = Z[f"{name}/flows"]
    randoms = Z[f"{name}/randoms"]
    layers = build_oracle_layers(Z, name)
    comp = C.CompositorOracle(flows.shape[1], flows.shape[2], layers, (0x20, 0x40, 0x60))
    ri = 0
    for t in range(flows.shape[0]):
        pm, rnd = {}, {}
        for li, layer in enumerate(layers):
            pm[li] = [G.pixmap_at(fr, t) for fr, _ in G.case_sources(Z, name, li)]
            if layer.spec.reset_mode == "random":
                rnd[li] = randoms[ri]
                ri += 1
        comp.update(flows[t], pm, rnd)
        out = comp.render()
        for li, layer in enumerate(layers):
            if layer.data is not None:
                np.testing.assert_array_equal(layer.data, Z[f"{name}/data{li}"][t], err_msg=f"{name} data t={t}")
        np.testing.assert_array_equal(out, Z[f"{name}/render"][t], err_msg=f"{name} render t={t}")


def test_reference_known_answers():
    """tests/test_compositor.py:29-54 of the reference."""
    k = G.load("kat_golden.npz")
    flow = k["flow"]
    lay = C.LayerOracle(C.LayerSpec(), 2, 3)
    lay.update(flow)
    assert tuple(lay.data[0, 0, :2]) == (1, 0) and tuple(lay.data[0, 1, :2]) == (1, 1)
    np.testing.assert_array_equal(lay.data, k["moveref"])
    lay = C.LayerOracle(C.LayerSpec(reset_mode="random", reset_random_factor=1), 2, 3)
    np.random.seed(5)
    lay.update(flow)
    assert tuple(lay.data[0, 0, :2]) == (0, 0) and tuple(lay.data[0, 1, :2]) == (0, 1)
    np.testing.assert_array_equal(lay.data, k["moveref_reset"])
    rm = np.zeros((2, 3), np.float32)
    rm[:, :1] = 1
    lay = C.LayerOracle(C.LayerSpec(reset_mode="random", reset_random_factor=1), 2, 3, reset_mask=rm)
    np.random.seed(5)
    lay.update(flow)
    assert tuple(lay.data[0, 0, :2]) == (0, 0) and tuple(lay.data[0, 1, :2]) == (1, 1)
    np.testing.assert_array_equal(lay.data, k["moveref_reset_mask"])


@pytest.mark.parametrize("params", [(0.5, 3, 15, 3, 5, 1.2), (0.5, 3, 19, 3, 7, 1.5), (0.7, 4, 15, 2, 5, 1.1)])
def test_farneback_restatement_matches_cv2(params):
    import cv2
    from transflow_b200.synthetic import synthetic_clip
    clip = synthetic_clip(120, 168, 2, seed=1)
    g = [cv2.cvtColor(f, cv2.COLOR_BGR2GRAY) for f in clip]
    ref = cv2.calcOpticalFlowFarneback(g[0], g[1], None, *params, 0)
    mine = FB.farneback(g[0], g[1], *params)
    epe = np.linalg.norm(mine - ref, axis=-1)
    assert epe.mean() < 1e-5 and epe.max() < 1e-4


FLOWZ = G.load("flow_golden.npz")


def replay_flow_source(method_cfg, direction, clip):
    """Re-run the reference's CvFlowSource.next + post_process chain with the oracle functions."""
    import json
    from oracle import flow_cv as F
    cfg = json.loads(method_cfg)
    method = cfg.pop("method")
    grays = [F.gray_from_bgr(f) for f in clip]
    prev_flow, out = None, []
    for t in range(1, len(grays)):
        left, right = (grays[t - 1], grays[t]) if direction == "forward" else (grays[t], grays[t - 1])
        if method == "farneback":
            raw = F.farneback(left, right, **{k[3:]: v for k, v in cfg.items()})
        elif method == "horn-schunck":
            raw = F.horn_schunck(left, right, None if prev_flow is None else prev_flow.copy(),
                                 cfg["hs_alpha"], cfg["hs_iterations"], cfg["hs_decay"], cfg["hs_delta"])
        else:
            raw = F.lucas_kanade(left, right, cfg["lk_window_size"], cfg["lk_max_level"], cfg["lk_step"])
        prev_flow = F.post_process(raw, direction == "forward")      # Q5: aliases the mutated array
        out.append(prev_flow.copy())
    return np.stack(out)


@pytest.mark.parametrize("name", sorted({k.split("/")[0] for k in FLOWZ.files if "/" in k}))
@pytest.mark.parametrize("direction", ["forward", "backward"])
def test_flow_oracle_matches_reference_flow_source(name, direction):
    import cv2
    clip = FLOWZ["clip"]
    from oracle import flow_cv as F
    for f in clip:
        np.testing.assert_array_equal(F.gray_from_bgr(f), cv2.cvtColor(f, cv2.COLOR_BGR2GRAY))
    got = replay_flow_source(str(FLOWZ[f"{name}/config"]), direction, clip)
    want = FLOWZ[f"{name}/{direction}"]
    if name.startswith("horn"):
        # float64/float32 evaluation-order noise only (2.4e-7)
        if direction == "backward":
            assert np.abs(got - want).max() < 1e-4
        else:
            # forward flows are integer-valued after the scatter; allow rare rounding flips
            assert (np.abs(got - want).max(axis=-1) > 0).mean() < 1e-3
    else:
        np.testing.assert_array_equal(got, want)


PPZ = G.load("postprocess_golden.npz")


@pytest.mark.parametrize("name", G.POSTPROCESS_CASES)
def test_postprocess_oracle_matches_reference(name):
    """Filters, mask and convolution kernel of ``FlowSource.post_process`` as the reference ran them
    (bit-exact; ``polar`` goes through libm transcendentals and gets 1e-5)."""
    import numpy  # the filter expressions of the golden cases name it, as transflow/utils.py does
    from oracle import flow_cv as F
    direction, filters, mask, kernel = G.postprocess_case_inputs(PPZ, name)
    grays = [F.gray_from_bgr(f) for f in PPZ["clip"]]
    want = PPZ[f"{name}/flows"]
    for i in range(1, len(grays)):
        left, right = (grays[i - 1], grays[i]) if direction == "forward" else (grays[i], grays[i - 1])
        t = i / float(PPZ["framerate"])      # output_frame_index is incremented before post_process (source.py:321)
        scalars = []
        for key, expr in filters:
            if key == "polar":
                scalars.append((key, tuple(eval("lambda t,r,a: " + e, {"numpy": numpy}) for e in expr)))
            else:
                scalars.append((key, eval("lambda t: " + expr, {"numpy": numpy})(t)))
        got = F.post_process(F.farneback(left, right), direction == "forward", mask=mask, kernel=kernel,
                             filters=scalars, t=t)
        assert got.dtype == want.dtype
        if any(k == "polar" for k, _ in filters):
            np.testing.assert_allclose(got, want[i - 1], atol=1e-5)
        else:
            np.testing.assert_array_equal(got, want[i - 1])


MERGE_MODES = ["first", "sum", "average", "difference", "product", "maskbin", "masklin", "absmax"]


@pytest.mark.parametrize("mode", MERGE_MODES)
def test_merge_oracle_matches_reference(mode):
    from oracle import flow_cv as F
    z = G.load("merge_golden.npz")
    for n in ((2,) if mode == "absmax" else (1, 2, 3)):
        got = F.merge_flows([z[f"flow{i}"].copy() for i in range(n)], mode)
        np.testing.assert_array_equal(got, z[f"{mode}/{n}"])
        assert got.dtype == z[f"{mode}/{n}"].dtype
    np.testing.assert_array_equal(F.upscale_array(z["flow0"].copy(), 3, 2), z["upscale/3x2"])


RENDER_CASES = {
    "2d/default": dict(kind="2d", scale=1, colors=None),
    "2d/scaled": dict(kind="2d", scale=0.37, colors=("#ff8000", "#0080ff", "rgb(12, 200, 77)", "#101010")),
    "2d/strong": dict(kind="2d", scale=3, colors=None),
    "1d/default": dict(kind="1d", scale=1, colors=None, binary=False),
    "1d/scaled": dict(kind="1d", scale=0.21, colors=("#203040", "#f0e0d1"), binary=False),
    "1d/binary": dict(kind="1d", scale=0.5, colors=("#ff0000", "#00ffff"), binary=True),
}


@pytest.mark.parametrize("name", sorted(RENDER_CASES))
def test_render_oracle_matches_reference(name):
    """oracle render1d / render2d vs frames the reference's output/render.py produced (render_golden.npz)."""
    from oracle import flow_cv as F
    z = G.load("render_golden.npz")
    c = RENDER_CASES[name]
    if c["kind"] == "2d":
        got = F.render2d(z["flow"], c["scale"], c["colors"])
    else:
        got = F.render1d(F.flow_magnitude(z["flow"]), c["scale"], c["colors"], c["binary"])
    assert got.dtype == np.uint8
    np.testing.assert_array_equal(got, z[name])


def test_flow_archive_writer_matches_reference_format(tmp_path):
    """Our NumpyOutput writes what the reference's NumpyOutput wrote (tests/golden/ref_archive.flow.zip): same
    member names, same meta.json, identical arrays."""
    import json
    import zipfile
    from transflow_b200.output import NumpyOutput
    from transflow_b200.output.zip import find_unique_path
    z = G.load("merge_golden.npz")
    ref = zipfile.ZipFile(os.path.join(G.GOLDEN_DIR, "ref_archive.flow.zip"))
    meta = json.loads(ref.read("meta.json").decode())
    path = str(tmp_path / "out.flow.zip")
    out = NumpyOutput(path, replace=True)
    out.write_meta(meta)
    for i in range(3):
        out.write_array(z[f"flow{i}"])
    out.close()
    mine = zipfile.ZipFile(path)
    assert sorted(mine.namelist()) == sorted(ref.namelist())
    assert json.loads(mine.read("meta.json").decode()) == meta
    for name in ref.namelist():
        if name.endswith(".npy"):
            np.testing.assert_array_equal(np.load(mine.open(name)), np.load(ref.open(name)))
    assert find_unique_path(path) == str(tmp_path / "out.000.flow.zip")
    open(tmp_path / "out.000.flow.zip", "w").close()
    assert find_unique_path(str(tmp_path / "out.000.flow.zip")) == str(tmp_path / "out.001.flow.zip")


def test_horn_schunck_own_blur_matches_cv2_blur():
    from oracle import flow_cv as F
    clip = FLOWZ["clip"]
    g0, g1 = F.gray_from_bgr(clip[0]), F.gray_from_bgr(clip[1])
    a = F.horn_schunck(g0, g1, use_blur_from_cv2=True)
    b = F.horn_schunck(g0, g1, use_blur_from_cv2=False)
    assert np.abs(a - b).max() < 1e-4


@pytest.mark.parametrize("case", [(96, 128, 1, 3), (120, 160, 4, 1), (67, 93, 2, 5)])
def test_lk_restatement_matches_cv2(case):
    from oracle import flow_cv as F, lk_np
    from transflow_b200.synthetic import synthetic_clip
    import cv2
    h, w, step, seed = case
    clip = synthetic_clip(h, w, 2, seed=seed)
    g0, g1 = F.gray_from_bgr(clip[0]), F.gray_from_bgr(clip[1])
    np.testing.assert_array_equal(lk_np.pyr_down(g0), cv2.pyrDown(g0))
    ref = F.lucas_kanade(g0, g1, 15, 2, step)
    mine = lk_np.dense_flow(g0, g1, 15, 2, step)
    d = np.abs(mine - ref).max(axis=-1)
    assert (d > 0).mean() < 1e-3 and d.max() < 1e-3


def test_philox_restatement_matches_random123_known_answers():
    """Random123's kat_vectors for philox4x32 (7 and 10 rounds): the restatement the device draws are checked with."""
    from oracle.philox_np import philox4x32, reset_draws
    kats = [
        (7, (0, 0, 0, 0), (0, 0), (0x5f6fb709, 0x0d893f64, 0x4f121f81, 0x4f730a48)),
        (7, (0xffffffff,) * 4, (0xffffffff,) * 2, (0x5207ddc2, 0x45165e59, 0x4d8ee751, 0x8c52f662)),
        (7, (0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
         (0x4dfccaba, 0x190a87f0, 0xc47362ba, 0xb6b5242a)),
        (10, (0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
        (10, (0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
    ]
    for rounds, ctr, key, want in kats:
        got = tuple(int(x) for x in philox4x32(ctr, key, rounds))
        assert got == want, (rounds, [hex(g) for g in got])
    u = reset_draws(0x5EED, 3, 33, 47)
    assert u.shape == (33, 47) and u.dtype == np.float64 and 0.0 <= u.min() and u.max() < 1.0
    assert abs(u.mean() - 0.5) < 0.03
    # a pixel pair shares one block; frames and seeds give unrelated fields
    assert not np.array_equal(u, reset_draws(0x5EED, 4, 33, 47)) and not np.array_equal(u, reset_draws(1, 3, 33, 47))


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_forward_claims_are_the_forward_flow(seed):
    """DESIGN.md 4a on the CPU: the claim plane of the forward scatter determines the forward flow (claimant - position,
    0 where unclaimed), and a move-reference layer fed with that flow fetches exactly the claimant's record -- so a layer
    that reads the claims needs no flow.  ``post_process`` is pinned to the reference by the golden flows above."""
    from oracle import flow_cv as F
    rng = np.random.default_rng(seed)
    h, w = 37, 53
    raw = rng.uniform(-7, 7, (h, w, 2)).astype(np.float32)
    raw[rng.random((h, w)) < 0.25] = 0
    mask = rng.random((h, w)).astype(np.float32) if seed == 1 else None
    filters = [("scale", 1.5)] if seed == 2 else []
    claims = F.forward_claims(raw, mask=mask, filters=filters)
    want = F.post_process(raw.copy(), True, mask=mask, filters=filters)
    ys, xs = np.mgrid[0:h, 0:w]
    claimant = np.where(claims > 0, claims - 1, ys * w + xs)
    got = np.stack([claimant % w - xs, claimant // w - ys], axis=-1).astype(np.float32)
    np.testing.assert_array_equal(got, np.asarray(want, np.float32))
    # what the compositor's move step fetches for that flow (movement.py:25-33: p + rint(fy) * w + rint(fx)) is the claimant
    off = np.rint(want[..., 1]).astype(np.int64) * w + np.rint(want[..., 0]).astype(np.int64)
    np.testing.assert_array_equal((ys * w + xs + off), claimant)
    assert ((claims > 0) & (claimant != ys * w + xs)).mean() > 0.3        # most pixels are claimed by another pixel
