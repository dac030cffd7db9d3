"""Zip container shared by the flow and checkpoint exports (``transflow/output/zip.py:6-29``)."""
import os


def find_unique_path(path: str) -> str:
    """First free name among ``path``, ``stem.000.ext``, ``stem.001.ext`` ... (``utils.py:147-160``): the double
    extensions ``.flow.zip`` / ``.map.*`` stay together, and an existing ``.NNN`` counter in the stem is continued."""
    import re
    stem, ext = os.path.splitext(path)
    for inner in (".flow", ".map"):
        if stem.endswith(inner):
            stem, ext = stem[:-len(inner)], inner + ext
            break
    counter = 0
    numbered = re.fullmatch(r".*\.(\d{3})", stem)
    if numbered:
        counter = int(numbered.group(1)) + 1
        stem = stem[:-4]
    while os.path.isfile(path):
        path = f"{stem}.{counter:03d}{ext}"
        counter += 1
    return path


class ZipOutput:

    def __init__(self, path: str, replace: bool = False):
        import zipfile
        self.path = path if replace else find_unique_path(path)
        if os.path.isfile(self.path):
            os.remove(self.path)
        self.archive = zipfile.ZipFile(self.path, "w", compression=zipfile.ZIP_DEFLATED)

    def write_meta(self, data: dict):
        if not data:
            return
        import json
        with self.archive.open("meta.json", "w") as file:
            file.write(json.dumps(data).encode())

    def write_object(self, filename: str, obj: object):
        import pickle
        with self.archive.open(filename, "w") as file:
            pickle.dump(obj, file)

    def close(self):
        self.archive.close()
