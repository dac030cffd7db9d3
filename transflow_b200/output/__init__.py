"""Outputs on the data path of the flow: ``.flow.zip`` export (``transflow/output/numpy.py``, ``zip.py``)."""
from .numpy import NumpyOutput
from .zip import ZipOutput

__all__ = ["NumpyOutput", "ZipOutput"]
