from . import diis  # noqa: F401
