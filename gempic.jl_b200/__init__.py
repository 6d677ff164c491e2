"""gempic.jl_b200 -- host-side mirror of GEMPIC.jl's particle-mesh API over libgempic_b200.so.

Only the north-star hot path lives here: ParticleGroup, ParticleMeshCoupling1D/2D,
Maxwell1DFEM, HamiltonianSplitting (1d1v, 1d2v), HamiltonianSplittingBoris and the
diagnostics that read particles every step.  Names, argument meaning and error behaviour
follow the reference (GEMPIC.jl/src); every call goes through the C ABI of
include/gempic_b200.h into hand-written sm_100a kernels.  No CPU fallback exists.

The directory name contains a dot, so it is loaded with __graft_entry__.load_package()
(module name `gempic_jl_b200`) rather than a plain `import`.
"""
from ._lib import ArgumentError, AssertionFailed, GempicError, device_count, finalize, init, init_devices, load  # noqa: F401
from .api import *  # noqa: F401,F403
from .api import __all__ as _api_all
from .dist import DistributedContext, shard_range  # noqa: F401
from .selfcheck import sharded_parity  # noqa: F401

__all__ = list(_api_all) + ["init", "init_devices", "device_count", "finalize", "load", "GempicError", "ArgumentError", "AssertionFailed",
                            "DistributedContext", "shard_range", "sharded_parity"]
