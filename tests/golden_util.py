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
