"""fasthevc_b200 - B200-native HEVC all-intra mode-decision engine (drop-in for the
TEncCu::xCompressCU hot path of omricarmi/FastHEVC = HM 16.14).

The product is the C-ABI library ``libfasthevc_cuda.so`` (include/fasthevc_cuda.h,
sources in fasthevc_b200/csrc: mode decision, deblocking, SAO statistics) and, next to
it, ``libfasthevc_cnn.so`` (include/fasthevc_cnn.h: the tcgen05 CU-split predictor) and
``libfasthevc_spool.so`` (the host encoder processes' view of the device's results).
This package is the thin Python host side used by tests and bench: ctypes mirrors of
those ABIs (api.py, cnn.py), the frame-group sharding (sharding.py), the host pipeline
(pipeline.py) and the dataset dump (harp_dump.py).  There is no CPU path: importing works
without a GPU, every compute call raises if the CUDA library or a device is missing.
"""
from .api import (Encoder, FhcError, SliceParams, SeqParams, CTU_RESULT_DTYPE, SAO_STATS_DTYPE, build_library, library_path, load_library,
                  pictures_in_flight, slice_params_from_qp, measure_int_issue)

__all__ = ["Encoder", "FhcError", "SliceParams", "SeqParams", "CTU_RESULT_DTYPE", "SAO_STATS_DTYPE", "build_library", "library_path",
           "load_library", "pictures_in_flight", "slice_params_from_qp", "measure_int_issue"]
