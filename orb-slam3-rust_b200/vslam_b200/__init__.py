"""vslam_b200 — B200-native stereo ORB front end behind the reference's StereoProcessor interface.

Python here is host plumbing over the C ABI (include/vslam_b200.h -> libvslam_b200.so).  The compute is
hand-written CUDA for sm_100a; there is no CPU or PyTorch fallback.
"""
from ._abi import DMATCH_DTYPE, KP_DTYPE, SEARCH_LOCAL_MAP, SEARCH_MOTION_MODEL, STAGE_BLURRED, STAGE_PYRAMID  # noqa: F401
from .stereo import (  # noqa: F401
    NN_RATIO,
    TH_HIGH,
    TH_LOW,
    CameraModel,
    FeatureGrid,
    FeatureSet,
    Frame,
    StereoFrame,
    StereoProcessor,
    VslamError,
    descriptor_distance,
)
