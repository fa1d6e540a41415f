"""repyolo_b200: B200-native (sm_100a) implementation of Rep-YOLO's deployed inference hot path.

Public surface = the reference's own call signatures for this path:
    Model(cfg).fuse().forward(x)      (models/yolo.py)      -> (pred [B, N, 5+nc], [raw heads])
    non_max_suppression(pred, ...)    (utils/general.py)    -> list of (n, 6) tensors
    letterbox / preprocess / scale_coords (utils/datasets.py, utils/general.py): the callers either side of the path
Everything numerical runs in csrc/librepyolo_b200.so (include/repyolo_b200.h); there is no CPU fallback.
"""
from .model import Model, IDetect, NativeEngine          # noqa: F401
from .nms import non_max_suppression, nms_padded, nms_launch_count          # noqa: F401
from .preproc import letterbox, preprocess, scale_coords, scale_detections  # noqa: F401
from ._lib import NativeError, lib                        # noqa: F401
from .arch import rep_yolo_cfg                            # noqa: F401
from .parallel import gather_detections, shard_bounds, to_list, DetectionGatherer  # noqa: F401
from .compat import attempt_load, from_reference, TracedModel, Ensemble, install as install_into_reference  # noqa: F401
