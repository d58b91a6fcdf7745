"""indelope_b200 -- B200-native (sm_100a) implementation of indelope's per-region calling path.

  indelope_b200.cuda   ctypes binding of libindelope_cuda.so (include/indelope_cuda.h): the hot path
  indelope_b200.host   ctypes binding of libindelope_host.so (include/indelope_host.h): sweep, packing, VCF text
  indelope_b200.api    host-side mirror of the reference's callsemble / main loop over the two libraries
"""
__version__ = "0.1.0"
