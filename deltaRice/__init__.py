"""deltaRice — import-compatible shim of the reference's Python package (reference setup.py:151,
src/h5.pyx): `import deltaRice.h5` keeps working in user scripts (reference README.md:65-91,
tests/test.py:1-2) and resolves to the B200 codec in `deltarice_b200`."""
