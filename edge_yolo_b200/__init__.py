"""edge_yolo_b200: B200-native (sm_100a) kernels for the EdgeLine-YOLO custom-operator path, behind the
reference's module / function interface.  See DESIGN.md and include/edgeline_b200.h."""
from ._lib import EdgelineError  # noqa: F401

__version__ = "0.1.0"
__all__ = ["EdgelineError", "ops", "modules", "model", "nms", "loss", "install"]
