"""Import-path shim: ``from mri_acoustic_model import build_acoustic_model`` (reference
scripts/run_mri_video_inference.py:120-128, scripts/export_predicted_mels.py:20,64-67) resolves to the
sm_100a drop-in when this directory is given as --mri-code-dir / --mri_code_dir."""
import os
import sys

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)

from mri2speech_b200.acoustic import (  # noqa: E402,F401
    BiLSTMSumMerge,
    EffNetV2B2Backbone,
    GlobalAvgPool,
    MRIAcousticModel,
    OTNLikeCNNBiLSTM,
    build_acoustic_model,
)
