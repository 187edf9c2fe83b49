"""fasthevc_b200 - B200-native HEVC all-intra mode-decision engine (drop-in for the
TEncCu::xCompressCU hot path of omricarmi/FastHEVC = HM 16.14).

The product is the C-ABI library ``libfasthevc_cuda.so`` (include/fasthevc_cuda.h,
sources in fasthevc_b200/csrc).  This package is the thin Python host side used by
tests and bench: a ctypes mirror of that ABI.  There is no CPU path: importing works
without a GPU, every compute call raises if the CUDA library or a device is missing.
"""
from .api import (Encoder, FhcError, SliceParams, SeqParams, CTU_RESULT_DTYPE, SAO_STATS_DTYPE, build_library, library_path, load_library,
                  pictures_in_flight, slice_params_from_qp, measure_int_issue)

__all__ = ["Encoder", "FhcError", "SliceParams", "SeqParams", "CTU_RESULT_DTYPE", "SAO_STATS_DTYPE", "build_library", "library_path",
           "load_library", "pictures_in_flight", "slice_params_from_qp", "measure_int_issue"]
