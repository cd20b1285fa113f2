"""B200-native active-selection scoring for nihalsid/deep-active-semantic-segmentation.

Only the selection-scoring hot path lives here: hand-written sm_100a CUDA kernels
behind a C ABI (include/das_b200.h, csrc/), a ctypes binding (_lib.py, ops.py) and a mirror of the
reference's `active_selection` selector classes (active_selection/) with unchanged method signatures.
"""
__version__ = "0.1.0"
