"""mri2speech_b200 -- B200 (sm_100a) implementation of the rtMRI -> mel -> waveform inference path.

Host-side mirror of the reference's module API over the C ABI in include/m2s.h:
  vocoder.Generator                  <-> reference models.Generator
  acoustic.build_acoustic_model      <-> reference mri2speech_code/mri_acoustic_model.build_acoustic_model
  pipeline.*                         <-> the glue in reference scripts/run_mri_video_inference.py
"""
__version__ = "0.1.0"
