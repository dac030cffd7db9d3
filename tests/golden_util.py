"""Helpers to replay the committed golden vectors (tests/golden/*.npz)."""
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    return np.load(os.path.join(GOLDEN_DIR, name), allow_pickle=False)


def compositor_case_names(z):
    return sorted({k.split("/")[0] for k in z.files})


def case_layers(z, name):
    """-> list of LayerConfig kwargs dicts (paths to mask PNGs are replaced by the arrays)."""
    return eval(str(z[f"{name}/layers"]))  # noqa: S307 - our own fixture


def case_sources(z, name, li):
    out, si = [], 0
    while f"{name}/pixmap{li}_{si}" in z.files:
        out.append((z[f"{name}/pixmap{li}_{si}"], z[f"{name}/intro{li}_{si}"]))
        si += 1
    return out


def pixmap_at(frames, t):
    return frames[min(t, len(frames) - 1)]


def postprocess_case_inputs(z, name):
    """-> (direction, filters as [(name, expression string | (str, str))], mask | None, kernel | None) of one case of
    ``postprocess_golden.npz`` (arguments as they were given to the reference's ``FlowSource.from_args``)."""
    import json
    args = json.loads(str(z[f"{name}/args"]))
    filters = []
    for item in (args.get("flow_filters") or "").split(";"):
        if not item.strip():
            continue
        key, val = item.split("=", 1)
        parts = tuple(val.strip().split(":"))
        filters.append((key.strip(), parts[0] if len(parts) == 1 else parts))
    mask = z["mask"] if args.get("mask_path") else None
    kernel = z["kernel/" + args["kernel_path"][:-4]] if args.get("kernel_path") else None
    return args["direction"], filters, mask, kernel


POSTPROCESS_CASES = ["scale", "threshold_fw", "clip", "chain_mask_fw", "strong", "polar", "polar_mid", "kernel_box3",
                     "kernel_box3_fw", "kernel_rand45_mask_fw", "kernel_row7"]
