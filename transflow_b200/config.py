"""Parameter contract of the hot path: the same field names and defaults as the reference's
``LayerConfig`` / ``PixmapSourceConfig`` (``transflow/config.py:11-158``), so configs and
checkpoints written by either side load in the other."""
from dataclasses import asdict, dataclass, field

from .utils import parse_timestamp

_TRUE_WORDS = ("1", "on", "o", "oui", "yes", "y")


def parse_bool_arg(arg, default: bool) -> bool:
    if arg is None:
        return default
    if isinstance(arg, str):
        return arg.lower().strip() in _TRUE_WORDS
    return bool(arg)


_LAYER_BOOL_DEFAULTS = {
    "transparent_pixels_can_move": False,
    "pixels_can_move_to_empty_spot": True,
    "pixels_can_move_to_filled_spot": True,
    "moving_pixels_leave_empty_spot": False,
    "reset_source": False,
    "introduce_pixels_on_empty_spots": True,
    "introduce_pixels_on_filled_spots": True,
    "introduce_moving_pixels": True,
    "introduce_unmoving_pixels": True,
    "introduce_once": False,
    "introduce_on_all_filled_spots": False,
    "introduce_on_all_empty_spots": False,
}


@dataclass
class LayerConfig:
    index: int
    classname: str | None = None
    mask_alpha: str | None = None
    mask_src: str | None = None
    mask_dst: str | None = None
    transparent_pixels_can_move: bool | str | None = None
    pixels_can_move_to_empty_spot: bool | str | None = None
    pixels_can_move_to_filled_spot: bool | str | None = None
    moving_pixels_leave_empty_spot: bool | str | None = None
    reset_mode: str | None = None
    reset_mask: str | None = None
    reset_random_factor: float | None = None
    reset_constant_step: float | None = None
    reset_linear_factor: float | None = None
    reset_source: bool | None = None
    introduce_pixels_on_empty_spots: bool | None = None
    introduce_pixels_on_filled_spots: bool | None = None
    introduce_moving_pixels: bool | None = None
    introduce_unmoving_pixels: bool | None = None
    introduce_once: bool | None = None
    introduce_on_all_filled_spots: bool | None = None
    introduce_on_all_empty_spots: bool | None = None

    def __post_init__(self):
        if self.classname is None:
            self.classname = "moveref"
        if self.reset_mode is None:
            self.reset_mode = "off"
        # defaults when constructed directly (the CLI passes 0.1 for the random factor, quirk Q14)
        if self.reset_random_factor is None:
            self.reset_random_factor = 1
        if self.reset_constant_step is None:
            self.reset_constant_step = 1
        if self.reset_linear_factor is None:
            self.reset_linear_factor = 0.1
        for name, default in _LAYER_BOOL_DEFAULTS.items():
            setattr(self, name, parse_bool_arg(getattr(self, name), default))

    @classmethod
    def fromdict(cls, d: dict):
        known = {k: d[k] for k in cls.__dataclass_fields__ if k in d and k != "index"}
        known.setdefault("classname", "reference")
        return cls(d["index"], **known)

    def todict(self) -> dict:
        return asdict(self)


@dataclass
class PixmapSourceConfig:
    path: str
    seek_time: float | str | None = None
    alteration_path: str | None = None
    introduction_path: str | None = None
    repeat: int | None = 1
    layers: list | None = None

    def __post_init__(self):
        self.seek_time = parse_timestamp(self.seek_time)
        if self.repeat is None:
            self.repeat = 1
        if self.layers is None:
            self.layers = [0]

    @classmethod
    def fromdict(cls, d: dict):
        return cls(d["path"], **{k: d.get(k) for k in ("seek_time", "alteration_path", "introduction_path",
                                                        "repeat", "layers") if k in d})

    def todict(self) -> dict:
        return asdict(self)
