"""``.flow.zip`` writer: the interface of ``transflow/output/numpy.py`` (``NumpyOutput(path, replace)``,
``write_array``), one ``%09d.npy`` member per flow.  Device flows are copied to the host here -- the one D2H copy
of the export path."""
import io

import numpy as np

from .zip import ZipOutput


class NumpyOutput(ZipOutput):

    def __init__(self, path: str, replace: bool = False):
        super().__init__(path, replace)
        self.index = 0

    def write_array(self, array):
        host = array.detach().cpu().numpy() if hasattr(array, "detach") else np.asarray(array)
        payload = io.BytesIO()
        np.save(payload, host, allow_pickle=False)
        self.write_bytes(f"{self.index:09d}.npy", payload.getvalue())
        self.index += 1
