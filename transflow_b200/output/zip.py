"""Zip containers of the exports (``.flow.zip``, ``.ckpt.zip``): the interface of ``transflow/output/zip.py``
(``ZipOutput(path, replace)``, ``write_meta``, ``write_object``, ``close``) over ``ZipFile.writestr``."""
import json
import os
import pickle
import re
import zipfile

_COUNTER = re.compile(r"^(?P<stem>.*)\.(?P<n>\d{3})$")
_INNER_EXTENSIONS = (".flow", ".map")


def find_unique_path(path: str) -> str:
    """``path`` if it is free, else the first free ``stem.NNN.ext`` (``utils.py:147-160``): ``.flow.zip`` /
    ``.map.*`` count as one extension and a counter already present in the name is continued, not nested."""
    if not os.path.isfile(path):
        return path
    stem, ext = os.path.splitext(path)
    inner = next((e for e in _INNER_EXTENSIONS if stem.endswith(e)), "")
    if inner:
        stem, ext = stem[:-len(inner)], inner + ext
    start = 0
    numbered = _COUNTER.match(stem)
    if numbered:
        stem, start = numbered.group("stem"), int(numbered.group("n")) + 1
    n = start
    while os.path.isfile(f"{stem}.{n:03d}{ext}"):
        n += 1
    return f"{stem}.{n:03d}{ext}"


class ZipOutput:
    """Deflated zip, written member by member; usable as a context manager."""

    def __init__(self, path: str, replace: bool = False):
        self.path = path if replace else find_unique_path(path)
        self.archive = zipfile.ZipFile(self.path, mode="w", compression=zipfile.ZIP_DEFLATED)  # "w" truncates

    def write_bytes(self, name: str, payload: bytes):
        self.archive.writestr(name, payload)

    def write_meta(self, data: dict):
        if data:
            self.write_bytes("meta.json", json.dumps(data).encode())

    def write_object(self, filename: str, obj: object):
        self.write_bytes(filename, pickle.dumps(obj))

    def close(self):
        self.archive.close()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
