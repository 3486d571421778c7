from . import base, blocks, cfl, dist  # noqa: F401
