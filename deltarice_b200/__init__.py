"""deltarice_b200 — B200-native Delta-Rice encode/decode (HDF5 filter 32025) hot path.

Host-side mirror of the reference's interface for this path:
  deltarice_b200.h5.H5FILTER / register_h5_filter   <- deltaRice.h5 (reference src/h5.pyx:27,55-61)
  deltarice_b200.DeltaRice                          <- batch ("chunk scheduler") front-end over the C-ABI
All arithmetic runs in libh5deltarice_b200.so (hand-written sm_100a CUDA); there is no CPU path.
"""
from ._lib import DeltaRiceError, LIB_PATH  # noqa: F401
from .codec import DeltaRice, chunk_offsets, parse_cd_values, chunk_bound_bytes  # noqa: F401

__version__ = "0.1.0"
