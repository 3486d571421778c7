"""cfl -- B200-native drop-in for the compatibility-scoring hot path of
appier/compatibility-family-learning: same ``cfl.layers`` / ``cfl.ops`` / ``cfl.models`` /
``cfl.utils`` call surface for that path, executed by hand-written sm_100a kernels behind a C ABI
(cfl/_lib/libcfl_b200.so, include/cfl_b200.h).  There is no CPU fallback."""
__all__ = ["layers", "ops", "models", "utils", "ranking", "variables"]
