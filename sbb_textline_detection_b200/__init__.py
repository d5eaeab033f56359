"""B200-native tiled CNN segmentation hot path of qurator-spk/sbb_textline_detection.

Public surface: ``textline_detector`` (drop-in for the reference class on the hot-path methods),
``SbbModel`` (GPU model handle, duck-types the Keras model ``do_prediction`` expects); around them
``compat.bind_reference`` (the reference's own ``run()`` on the GPU methods), ``pipeline.PageDispatcher`` (many pages
through the three resident models), ``parallel`` (page-per-GPU sharding, weight broadcast, one page across GPUs),
``precision`` (per-layer precision plan), ``cli`` / ``ocrd_cli`` (the reference's console scripts)."""
from .model import SbbModel, SbbSession, compute_tile_grid  # noqa: F401


def __getattr__(name):  # detector imports cv2; keep `import sbb_textline_detection_b200` light
    if name == "textline_detector":
        from .detector import textline_detector
        return textline_detector
    raise AttributeError(name)
