"""GPU tests of the host-side plugin logic around the kernels: locked flows / Horn-Schunck decay with a mask or a
kernel (``prev_flow`` aliasing, reference ``source.py:293-363``), the documented filter expressions, and the
``.ckpt.zip`` write -> resume round trip (reference ``pipeline.py:225-242, 290-303``, ``tests/test_pipeline.py:90-119``)."""
import itertools
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

from tests import golden_util as G  # noqa: E402

LOCK_CASES = ["stay_mask_fw", "stay_kernel_bw", "stay_plain_fw", "skip_mask_bw", "hs_decay_mask_fw",
              "hs_decay_kernel_bw", "usage_math", "usage_polar_numpy"]


def write_avi(path, clip, fps=25):
    import cv2
    vw = cv2.VideoWriter(str(path), cv2.VideoWriter_fourcc(*"FFV1"), fps, (clip.shape[2], clip.shape[1]))
    assert vw.isOpened(), "cv2 cannot write FFV1"
    for f in clip:
        vw.write(f)
    vw.release()


@pytest.mark.parametrize("name", LOCK_CASES)
def test_flow_source_lock_and_decay_match_reference(name, tmp_path):
    """Flows the REAL reference produced (tests/golden/make_golden.py::lock_cases) for sources that read ``prev_flow``
    again.  A source that post-processed ``prev_flow`` fully in place next to a mask / kernel would re-apply them on
    every locked frame (the held flow decays) -- far outside these tolerances."""
    import PIL.Image
    from transflow_b200.flow import FlowSource
    from transflow_b200.flow.sources.cv import ArrayCapture
    z = G.load("lock_golden.npz")
    args = json.loads(str(z[f"{name}/args"]))
    cfg_path = tmp_path / "cfg.json"
    cfg_path.write_text(str(z["config/" + args.pop("cfg")]))
    if args.get("mask_path"):
        path = str(tmp_path / "mask.png")
        PIL.Image.fromarray(np.rint(z["mask"] * 255).astype(np.uint8)).save(path)
        args["mask_path"] = path
    if args.get("kernel_path"):
        path = str(tmp_path / "kernel.npy")
        np.save(path, z["kernel/box3"])
        args["kernel_path"] = path
    want = z[f"{name}/flows"]
    with FlowSource.from_args(ArrayCapture(z["clip"], 25.0), cv_config=str(cfg_path), **args) as src:
        flows = list(itertools.islice(src, len(want)))
    assert len(flows) == len(want)
    for t, (got, ref) in enumerate(zip(flows, want)):
        assert got.dtype == np.float32 and got.shape == ref.shape
        err = np.linalg.norm(got.astype(np.float64) - ref, axis=-1)
        if args["direction"] == "backward":
            assert err.mean() <= 0.01 and err.max() <= 0.1, (name, t, err.mean(), err.max())
        else:       # forward: integer targets; a raw flow 1e-6 off can flip a rounding on a few pixels
            assert (err > 0).mean() < 1e-2, (name, t, (err > 0).mean())


def test_filter_expressions_see_math_random_numpy():
    """USAGE.md documents ``math`` / ``random`` / ``numpy`` inside expressions (reference utils.py:1-8, :409-414)."""
    from transflow_b200.utils import parse_lambda_expression
    assert abs(parse_lambda_expression("1-math.exp(-.5*t)")(2.0) - 0.6321205588285577) < 1e-15
    assert 0.0 <= parse_lambda_expression("random.random()*t")(1.0) <= 1.0
    assert parse_lambda_expression("numpy.sqrt(t)")(4.0) == 2.0


def _run_pipeline(cfg, **kw):
    from transflow_b200.pipeline import Pipeline
    pipe = Pipeline(cfg, **kw)
    pipe.run()
    return pipe


@pytest.mark.parametrize("layer_kw", [dict(), dict(reset_mode="random", reset_random_factor=0.5)])
def test_checkpoint_file_round_trip(tmp_path, layer_kw):
    """The reference's ``test_checkpoint``: run ckpt + 1 frames writing a checkpoint every ckpt frames, resume from
    the FILE, and the resumed last frame equals the uninterrupted one (diff == 0)."""
    import PIL.Image
    from transflow_b200.config import LayerConfig, PixmapSourceConfig
    from transflow_b200.pipeline import Config
    from transflow_b200.synthetic import synthetic_clip
    ckpt, fps = 5, 25
    h, w = 72, 96
    clip = synthetic_clip(h, w, ckpt + 3, seed=4)
    avi = tmp_path / "flow.avi"
    write_avi(avi, clip, fps)
    out1 = tmp_path / "1-%d.png"
    cfg = Config(str(avi), pixmap_sources=[PixmapSourceConfig("cnoise", layers=[0])],
                 layers=[LayerConfig(0, "moveref", **layer_kw)], output_path=str(out1), direction="forward",
                 duration_time=(ckpt + 1) / fps, seed=7)
    pipe = _run_pipeline(cfg, checkpoint_every=ckpt)
    assert pipe.cursor == ckpt + 1 and pipe.expected_length == ckpt + 1
    ckpt_path = tmp_path / f"1-%d_{ckpt:05d}.ckpt.zip"
    assert ckpt_path.is_file(), sorted(p.name for p in tmp_path.iterdir())
    for i in range(ckpt + 1):
        assert (tmp_path / f"1-{i}.png").is_file()
        os.rename(tmp_path / f"1-{i}.png", tmp_path / f"2-{i}.png")
    pipe2 = _run_pipeline(Config(str(ckpt_path)), checkpoint_every=ckpt)
    assert pipe2.cursor == ckpt + 1 and pipe2.expected_length == 1
    assert len(list(tmp_path.glob("1-*.png"))) == 1
    a = np.asarray(PIL.Image.open(tmp_path / f"1-{ckpt}.png"))
    b = np.asarray(PIL.Image.open(tmp_path / f"2-{ckpt}.png"))
    assert a.shape == (h, w, 3) and np.array_equal(a, b)
    # the archive holds what the reference writes (pipeline.py:227-233)
    import zipfile
    with zipfile.ZipFile(ckpt_path) as z:
        assert sorted(z.namelist()) == ["compositor.bin", "meta.json"]
        meta = json.loads(z.read("meta.json"))
    assert meta["cursor"] == ckpt and meta["framerate"] == fps and meta["config"]["flow_path"] == str(avi)


@pytest.mark.parametrize("layer_kw", [dict(), dict(reset_mode="random", reset_random_factor=0.5)])
def test_pipeline_forward_claims_equal_flow_path(tmp_path, monkeypatch, layer_kw):
    """A forward pipeline whose only flow consumer is a single move-reference layer hands the scatter pass's claim plane
    to the compositor (``flow_source.output == "claims"``); the frames equal those of the flow-fed path bit for bit, and
    a pipeline with a flow visualiser keeps the flow."""
    import PIL.Image
    from transflow_b200.config import LayerConfig, PixmapSourceConfig
    from transflow_b200.pipeline import Config
    from transflow_b200.synthetic import synthetic_clip
    h, w, n = 72, 96, 6
    avi = tmp_path / "flow.avi"
    write_avi(avi, synthetic_clip(h, w, n + 2, seed=9), 25)
    frames = {}
    for tag, env in (("claims", "1"), ("flow", "0")):
        monkeypatch.setenv("TFB200_FORWARD_CLAIMS", env)
        cfg = Config(str(avi), pixmap_sources=[PixmapSourceConfig("cnoise", layers=[0])],
                     layers=[LayerConfig(0, "moveref", **layer_kw)], output_path=str(tmp_path / (tag + "-%d.png")),
                     direction="forward", duration_time=n / 25, seed=3)
        pipe = _run_pipeline(cfg)
        assert pipe.flow_source.output == ("claims" if env == "1" else "device")
        frames[tag] = [np.asarray(PIL.Image.open(tmp_path / f"{tag}-{i}.png")) for i in range(n)]
    for i in range(n):
        assert np.array_equal(frames["claims"][i], frames["flow"][i]), i
    assert any(not np.array_equal(frames["claims"][0], f) for f in frames["claims"][1:])   # the frames do change
    monkeypatch.setenv("TFB200_FORWARD_CLAIMS", "1")
    cfg = Config(str(avi), pixmap_sources=[PixmapSourceConfig("cnoise", layers=[0])], layers=[LayerConfig(0, "moveref")],
                 output_path=str(tmp_path / "v-%d.png"), direction="forward", duration_time=2 / 25, view_flow=True)
    assert _run_pipeline(cfg).flow_source.output == "device"


def test_layer_masks_and_seed_survive_pickle():
    """A ``random`` mask is redrawn from its config string on construction; the pickled layer must carry the array
    (the reference pickles its mask arrays).  Two layers of one compositor draw different reset numbers."""
    import pickle
    from transflow_b200.compositor import Compositor
    from transflow_b200.config import LayerConfig
    h, w = 24, 40
    comp = Compositor.from_args(h, w, [LayerConfig(0, "moveref", reset_mode="random", reset_mask="random"),
                                       LayerConfig(1, "moveref", reset_mode="random")], seed=11)
    assert comp.layers[0].rng_seed != comp.layers[1].rng_seed
    other = Compositor.from_args(h, w, [LayerConfig(0, "moveref", reset_mode="random")], seed=12)
    assert other.layers[0].rng_seed != comp.layers[0].rng_seed
    clone = pickle.loads(pickle.dumps(comp))
    np.testing.assert_array_equal(clone.layers[0].reset_mask, comp.layers[0].reset_mask)
    assert clone.layers[0].rng_seed == comp.layers[0].rng_seed


def test_level_a_reference_pipeline_runs_with_swapped_classes(tmp_path):
    """INTEGRATION.md level A: the reference's own multi-process ``Pipeline`` (flow SourceProcess -> queue ->
    main loop -> output process, reference pipeline.py:56-136, 440-455, 545-575) drives this package's
    ``FlowSource`` / ``Compositor`` / ``PixmapSourceInterface`` in place of its own, and writes the frames the
    stock run writes (backward direction: integer state bit-exact given flows within 1e-5 px of cv2)."""
    import subprocess
    import sys
    import PIL.Image
    from transflow_b200.synthetic import synthetic_clip
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    if not os.path.isfile(os.path.join(root, "baseline", "_ref", "transflow", "pipeline.py")):
        pytest.skip("baseline/_ref (the installed reference) is not present")
    h, w, n = 96, 128, 6
    avi = tmp_path / "flow.avi"
    write_avi(avi, synthetic_clip(h, w, n, seed=9))
    outs = {}
    for mode in ("stock", "swapped"):
        out = tmp_path / mode
        env = dict(os.environ, PYTHONDONTWRITEBYTECODE="1")
        if mode == "stock":
            env["CUDA_VISIBLE_DEVICES"] = ""        # the stock run is the reference's CPU path, nothing else
        res = subprocess.run([sys.executable, os.path.join(root, "tests", "level_a_runner.py"), mode, str(avi), str(out)],
                             capture_output=True, text=True, timeout=600, env=env)
        assert res.returncode == 0 and f"cursor {n - 1}" in res.stdout, (mode, res.stdout[-2000:], res.stderr[-4000:])
        if mode == "swapped":
            assert "transflow_b200.compositor" in res.stdout, res.stdout
        frames = sorted(out.glob("*.png"), key=lambda p: int(p.stem))
        assert len(frames) == n - 1, (mode, [p.name for p in frames])
        outs[mode] = [np.asarray(PIL.Image.open(p)) for p in frames]
    for t, (a, b) in enumerate(zip(outs["stock"], outs["swapped"])):
        assert a.shape == (h, w, 3)
        np.testing.assert_array_equal(a, b, err_msg=f"frame {t}")


def test_plugin_errors_are_the_reference_exceptions(tmp_path):
    """SURVEY.md 8(b) error contract: the exceptions the reference raises at the same places (cv.py:474-476, 518;
    source.py:313, 294-295; layer.py:55; NumPy's IndexError for a flow that points outside the frame)."""
    from transflow_b200 import ops
    from transflow_b200.compositor import Compositor
    from transflow_b200.compositor.pixmap_source_interface import PixmapSourceInterface, StillQueue
    from transflow_b200.config import LayerConfig
    from transflow_b200.flow import FlowSource
    from transflow_b200.flow.sources.cv import ArrayCapture, CvFlowConfig, CvFlowSource
    from transflow_b200.synthetic import synthetic_clip
    h, w = 48, 64
    clip = synthetic_clip(h, w, 3, seed=1)
    with pytest.raises(ValueError):
        CvFlowSource.Method.from_string("block-matching")
    with pytest.raises(ValueError):
        FlowSource.Direction.from_arg("sideways")
    with pytest.raises(TypeError):
        CvFlowConfig(fb_window=15)                               # unknown field
    with pytest.raises(ImportError):                             # LiteFlowNet is outside the accelerated path
        with FlowSource.from_args(ArrayCapture(clip), cv_config=CvFlowConfig(method="liteflownet")) as src:
            next(src)
    with pytest.raises(RuntimeError):                            # locked before any flow exists (source.py:313)
        with FlowSource.from_args(ArrayCapture(clip), lock_expr="(0,1),(100,1)", lock_mode="stay") as src:
            next(src)
    with FlowSource.from_args(ArrayCapture(clip), direction="backward") as src:
        assert len(list(src)) == 2
        with pytest.raises(StopIteration):
            next(src)
    with pytest.raises(ValueError):
        Compositor.from_args(h, w, [LayerConfig(0, "hologram")])
    with pytest.raises(ValueError):
        ops.Farneback(h, w)(torch.zeros((h, w + 1), dtype=torch.uint8, device="cuda"),
                            torch.zeros((h, w), dtype=torch.uint8, device="cuda"))
    comp = Compositor.from_args(h, w, [LayerConfig(0, "moveref")])
    comp.set_sources({0: [PixmapSourceInterface(StillQueue(np.zeros((h, w, 3), np.uint8)), np.ones((h, w), bool))]})
    with pytest.raises(ValueError):
        comp.step(np.zeros((h, w + 2, 2), np.float32))           # flow of another size
    flow = np.zeros((h, w, 2), np.float32)
    flow[h - 1, w - 1] = (3, 2)                                   # un-clipped: points past the last pixel
    comp.step(flow)
    with pytest.raises(IndexError):                              # NumPy's .flat[] would have raised in the reference
        comp.layers[0].check_indices()
    bad = Compositor.from_args(h, w, [LayerConfig(0, "moveref")])
    bad.set_sources({0: [PixmapSourceInterface(StillQueue(np.zeros((h + 1, w, 3), np.uint8)), np.ones((h, w), bool))]})
    with pytest.raises(ValueError):
        bad.step(np.zeros((h, w, 2), np.float32))                # pixmap of another size


@pytest.mark.parametrize("src,dst", [((480, 854), (270, 480)), ((270, 480), (1080, 1920)), ((123, 457), (77, 301)),
                                      ((96, 128), (96, 128)), ((2160, 3840), (1080, 1920))])
def test_resize_nearest_bit_exact(src, dst):
    """cv2.resize(frame, (w, h), interpolation=INTER_NEAREST) as the reference calls it (cv.py:464), on the device."""
    import cv2
    from transflow_b200 import ops
    rng = np.random.default_rng(31)
    frame = rng.integers(0, 256, (*src, 3), dtype=np.uint8)
    got = ops.resize_nearest_bgr(torch.from_numpy(frame).cuda(), dst[0], dst[1]).cpu().numpy()
    np.testing.assert_array_equal(got, cv2.resize(frame, dsize=(dst[1], dst[0]), interpolation=cv2.INTER_NEAREST))


def test_flow_source_resizes_frames_of_another_size_on_device():
    """A capture that delivers frames of another size than it reports (a webcam ignoring set(), cv.py:455-457): the
    source resizes with INTER_NEAREST like the reference, here after the upload."""
    import cv2
    from oracle import flow_cv as F
    from transflow_b200.flow import FlowSource
    from transflow_b200.flow.sources.cv import ArrayCapture
    from transflow_b200.synthetic import synthetic_clip
    big = synthetic_clip(192, 256, 3, seed=8)

    class Lying(ArrayCapture):
        def get(self, prop):
            if prop == cv2.CAP_PROP_FRAME_WIDTH:
                return 128
            if prop == cv2.CAP_PROP_FRAME_HEIGHT:
                return 96
            return super().get(prop)

    with FlowSource.from_args(Lying(big), direction="backward") as src:
        assert (src.width, src.height) == (128, 96)
        flows = list(src)
    small = [cv2.resize(f, dsize=(128, 96), interpolation=cv2.INTER_NEAREST) for f in big]
    grays = [F.gray_from_bgr(f) for f in small]
    assert len(flows) == 2
    for t, got in enumerate(flows):
        want = F.post_process(F.farneback(grays[t + 1], grays[t]), False)
        err = np.linalg.norm(got - want, axis=-1)
        assert err.mean() <= 0.01 and err.max() <= 0.1, (t, err.mean(), err.max())
