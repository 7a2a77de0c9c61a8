"""recombiner_b200 -- B200-native (sm_100a) implementation of RECOMBINER's hot path:
batched per-datapoint variational INR fitting and the REC candidate search, behind
the reference's module API.  The compute lives in librecombiner_b200.so (C ABI in
include/recombiner_b200.h); this package is the Python host mirror."""
__all__ = ["config", "utils", "prior_model", "test_model", "engine", "rec"]
