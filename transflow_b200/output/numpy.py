"""``.flow.zip`` writer (``transflow/output/numpy.py:6-15``): one ``NNNNNNNNN.npy`` per frame."""
import numpy as np

from .zip import ZipOutput


class NumpyOutput(ZipOutput):

    def __init__(self, path: str, replace: bool = False):
        ZipOutput.__init__(self, path, replace)
        self.index = 0

    def write_array(self, array):
        if hasattr(array, "is_cuda"):      # a device flow: one D2H copy at this legacy edge
            array = array.cpu().numpy()
        with self.archive.open(f"{self.index:09d}.npy", "w") as file:
            np.save(file, array)
        self.index += 1
