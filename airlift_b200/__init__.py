"""airlift_b200 -- B200-native re-alignment hot path of AirLift's minimap2 fork.

The product is the C library (libmm2b200.so: CUDA kernels for sm_100a + C host mapper) and the
minimap2-compatible CLI built from airlift_b200/host; this Python package only binds the C-ABI
for the tests and bench.py."""
