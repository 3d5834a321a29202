"""deltaRice.h5 — the reference module's names (src/h5.pyx:27, :55-61) on top of deltarice_b200.h5:

    H5FILTER              filter id 32025
    register_h5_filter()  raises RuntimeError when registration fails, as the reference does

Importing the module registers the filter with the libhdf5 that h5py has loaded (the reference does the
same at import, src/h5.pyx:61).  Without h5py in the process nothing is registered and nothing is raised:
the filter is then only reachable through HDF5_PLUGIN_PATH or deltarice_b200.h5.apply_filter."""
from deltarice_b200.h5 import (H5FILTER, apply_filter, decode_dataset_chunks, encode_dataset_chunks,  # noqa: F401
                               plugin_path)
from deltarice_b200 import h5 as _h5


def register_h5_filter():
    ret = _h5.register_h5_filter()
    if ret < 0:
        raise RuntimeError("Failed to register DeltaRice HDF5 filter.", ret)


if _h5.h5py_loaded():
    register_h5_filter()
