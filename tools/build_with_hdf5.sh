#!/bin/sh
# SURVEY §8 f1 — the integration proof against a REAL libhdf5 (this image has none: the script says so
# and exits 0; run it on a box that has libhdf5 headers + h5py + a B200).
#   1. builds the plugin against the system's hdf5.h / H5PLextern.h (-DDRICE_USE_SYSTEM_HDF5) instead
#      of the vendored ABI slice in include/hdf5_abi/;
#   2. runs the reference's own tests/test.py UNMODIFIED (it does `import deltaRice.h5`: the shim
#      package at the repo root) and, separately, with plugin discovery only (HDF5_PLUGIN_PATH);
#   3. compiles and runs the reference's examples/testCode.c UNMODIFIED against include/deltaRice.h
#      and this library.
# usage: tools/build_with_hdf5.sh [/path/to/reference]        (default /root/reference)
set -e
HERE=$(cd "$(dirname "$0")/.." && pwd)
REF=${1:-/root/reference}
if command -v h5cc >/dev/null 2>&1; then
    H5CFLAGS=$(h5cc -show | tr ' ' '\n' | grep '^-I' | tr '\n' ' ')
    H5LIBS=$(h5cc -show | tr ' ' '\n' | grep -E '^-L|^-l|^-Wl' | tr '\n' ' ')
elif pkg-config --exists hdf5 2>/dev/null; then
    H5CFLAGS=$(pkg-config --cflags hdf5)
    H5LIBS=$(pkg-config --libs hdf5)
else
    echo "build_with_hdf5: no libhdf5 on this machine (neither h5cc nor pkg-config hdf5): nothing to do"
    exit 0
fi
echo "== 1. plugin against the system HDF5 headers"
make -C "$HERE/deltarice_b200/csrc" clean
make -C "$HERE/deltarice_b200/csrc" -j8 EXTRA="-DDRICE_USE_SYSTEM_HDF5 $H5CFLAGS"
export HDF5_PLUGIN_PATH="$HERE/deltarice_b200"
export PYTHONPATH="$HERE:$PYTHONPATH"
if python -c 'import h5py' 2>/dev/null; then
    echo "== 2a. reference tests/test.py, unmodified (import deltaRice.h5 -> shim)"
    (cd "$(mktemp -d)" && python -m pytest -q "$REF/tests/test.py")
    echo "== 2b. this repo's h5py tests (plugin discovery, stored chunks vs oracle)"
    python -m pytest -q -m gpu "$HERE/tests/test_real_hdf5.py"
else
    echo "== 2. skipped: no h5py"
fi
echo "== 3. reference examples/testCode.c, unmodified"
T=$(mktemp -d)
gcc -O2 -I"$HERE/include" -DDRICE_USE_SYSTEM_HDF5 $H5CFLAGS "$REF/examples/testCode.c" -o "$T/testCode" \
    -L"$HERE/deltarice_b200" -lh5deltarice_b200 $H5LIBS -Wl,-rpath,"$HERE/deltarice_b200"
(cd "$T" && timeout 600 ./testCode) || echo "testCode exited non-zero (its own loop counter overflows: SURVEY Appendix B12)"
echo "build_with_hdf5: done"
