"""CPU oracle for the rtMRI -> mel -> waveform inference path.

TEST INFRASTRUCTURE ONLY.  Nothing in the product path (``mri2speech_b200``,
``models.py``, ``mri2speech_code/``, ``scripts/``) may import this package; only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs do, and only as the checker or the timed CPU arm.

Pinning status (see DESIGN.md "Oracle"):
  * vocoder (``oracle.vocoder``): PINNED.  The restatement is checked in this
    container against the reference's own ``models.Generator`` (imported from
    /root/reference with a matplotlib stub, ``oracle/ref_import.py``); the
    generating script ``oracle/gen_golden.py`` commits small input/output
    vectors under ``tests/golden/``.
  * BiLSTM + head + mel glue (``oracle.acoustic``, ``oracle.glue``): restated on
    ``torch.nn.functional`` / explicit gate arithmetic and cross-checked against
    ``torch.nn.LSTM`` -- the exact call the reference makes
    (mri2speech_code/mri_acoustic_model.py:57-71).
  * frame-CNN encoder (``oracle.acoustic.encoder_forward``): PARITY UNPINNED.
    The arithmetic lives in the un-vendored dependency ``timm==1.0.21``
    (model id ``tf_efficientnetv2_b2``; requirements.lab.txt:11), absent from
    /root/reference and from this image.  We restate its published topology and
    anchor on the call site (mri_acoustic_model.py:28-48) and structural checks
    (8 391 406 backbone parameters, (N,208,8,8) output).
"""
