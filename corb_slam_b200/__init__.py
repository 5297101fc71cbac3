"""corb_slam_b200 - B200-native (sm_100a) hot path of CORB-SLAM behind the reference's own interfaces.

Only what the path needs: csrc/ (CUDA kernels + the C ABI of libcorb_b200.so) and host-side mirrors of the
reference classes that own the path (ORBextractor, ORBmatcher, ORBVocabulary, PnPsolver, Optimizer).
"""
from ._lib import CorbError, KP_DTYPE, LIB_PATH  # noqa: F401
from .orbextractor import (ORBextractor, compute_stereo_matches, extract_stereo, extract_stereo_device,  # noqa: F401
                           extract_stereo_submit, extract_stereo_wait, frame_stereo, frame_stereo_submit, frame_stereo_wait)
from .orbmatcher import ORBmatcher, BowFeatures  # noqa: F401
from .orbvocabulary import BowRecord, ORBVocabulary  # noqa: F401
from .frame import FrameView  # noqa: F401
from .pnpsolver import PnPsolver  # noqa: F401
from .optimizer import Optimizer, page_locked, torch_allreduce  # noqa: F401
