#!/usr/bin/env python
"""INTEGRATION.md level A, executed: the REFERENCE's own multi-process ``transflow.pipeline.Pipeline`` (installed
under baseline/_ref) with the three imports its pipeline.py makes for the hot path swapped for this package's
classes (``FlowSource``, ``Compositor``, ``PixmapSourceInterface``; reference pipeline.py:22-24), or left alone
(``stock``).  Run as a script by tests/test_pipeline_gpu.py -- the pipeline forks its source processes, so it wants a
fresh interpreter that has not touched CUDA.

    python tests/level_a_runner.py {stock|swapped} <flow video> <output dir> [direction]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.dont_write_bytecode = True
sys.path.insert(0, os.path.join(ROOT, "baseline", "_ref"))
sys.path.insert(1, ROOT)


def main():
    mode, video, outdir = sys.argv[1:4]
    direction = sys.argv[4] if len(sys.argv) > 4 else "backward"
    import transflow.pipeline as P
    from transflow.config import Config, LayerConfig, PixmapSourceConfig
    if mode == "swapped":
        from transflow_b200.compositor import Compositor
        from transflow_b200.compositor.pixmap_source_interface import PixmapSourceInterface
        from transflow_b200.flow import FlowSource
        P.FlowSource, P.Compositor, P.PixmapSourceInterface = FlowSource, Compositor, PixmapSourceInterface
    cfg = Config(video, pixmap_sources=[PixmapSourceConfig("cnoise", layers=[0])],
                 layers=[LayerConfig(0, "moveref")], output_path=os.path.join(outdir, "%d.png"),
                 direction=direction, seed=3)
    pipe = P.Pipeline(cfg, export_config=False, replace=True, log_handler="null")
    pipe.run()
    print("cursor", pipe.cursor, type(pipe.compositor).__module__)


if __name__ == "__main__":
    main()
