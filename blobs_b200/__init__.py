"""blobs_b200 — B200-native implementation of the `blobs::Physics::step` hot path.

`World` is the raw handle over the C ABI (include/blobs_b200.h); `blobs_b200.physics` mirrors the
reference's Rust API names (Physics, RigidBodyBuilder, ColliderBuilder, ...). Importing this package
loads libblobs_b200.so and fails loudly if it has not been built; there is no CPU fallback.
"""
from . import _abi as abi  # noqa: F401
from ._lib import LIB_PATH, load  # noqa: F401
from .world import BlobsError, World  # noqa: F401

load()
