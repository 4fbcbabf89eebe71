"""ser_b200: B200-native fast-profile acoustic front-end for jsugg/ser (hot path only).

The public surface mirrors the reference's fast-profile seams:

* ``ser_b200.dsp.extract_feature_from_signal``      <- ser/_internal/utils/dsp.py:67
* ``ser_b200.handcrafted.HandcraftedBackend``        <- ser/_internal/repr/handcrafted.py:22
* ``ser_b200.feature_extractor.extract_feature_frames`` <- ser/_internal/features/feature_extractor.py:164
* ``ser_b200.fast_inference.run_fast_inference``     <- ser/_internal/runtime/fast_inference.py:35

All arithmetic runs in hand-written sm_100a CUDA kernels behind the C-ABI library declared
in ``include/ser_b200.h``.  There is no CPU fallback: every compute entry point raises if
the library is missing or no CUDA device is present.
"""

__version__ = "0.1.0"
