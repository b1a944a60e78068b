"""imageanalysis_b200 — B200-native (sm_100a) drop-in for the pairwise
feature-matching hot path of NorthStarUAS/ImageAnalysis
(scripts/lib/matcher.py driven by scripts/process.py / 3a-matching.py).

  matcher   : the reference's `lib.matcher` API, served by libiamatch.so
  _capi     : ctypes binding of the C-ABI (include/iamatch.h)
  pairs     : pair work-list generators (matcher.py:858-903)
  dist      : pair-list sharding + NCCL all-gather of match tables
  synth     : synthetic inputs of the benchmark shapes
  build     : nvcc build of libiamatch.so for sm_100a
"""
__version__ = "0.1.0"
