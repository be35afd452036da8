"""Import-path shim: ``from models import Generator`` (reference scripts/run_mri_video_inference.py:19,
mel_to_audio_synthesis.py) resolves to the sm_100a drop-in."""
from mri2speech_b200.vocoder import Generator, ResBlock1, LRELU_SLOPE, get_padding, init_weights  # noqa: F401
