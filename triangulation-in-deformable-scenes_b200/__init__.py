"""B200-native deformable two-view hot path: batched triangulation + non-rigid LM refinement.

The product is csrc/ (hand-written sm_100a CUDA behind the C ABI of include/dsc.h, built into
lib/libdsc_b200.so) and host/ (the C++ mirror of the reference's Map / KeyFrame / MapPoint and
optimisation-call API).  `binding` is the ctypes face used by tests/ and bench.py.
The directory name is not a Python identifier; load it with `__graft_entry__.package()`.
"""
from . import binding  # noqa: F401
from .binding import Context, Batch, BundleAdjuster, pin_host, unpin_host, make_pair, make_weights, load_library, DscError, shard_partition  # noqa: F401
