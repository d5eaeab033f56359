"""B200-native tiled CNN segmentation hot path of qurator-spk/sbb_textline_detection.

Public surface: ``textline_detector`` (drop-in for the reference class on the hot-path methods),
``SbbModel`` (GPU model handle, duck-types the Keras model ``do_prediction`` expects)."""
from .model import SbbModel, SbbSession, compute_tile_grid  # noqa: F401


def __getattr__(name):  # detector imports cv2; keep `import sbb_textline_detection_b200` light
    if name == "textline_detector":
        from .detector import textline_detector
        return textline_detector
    raise AttributeError(name)
