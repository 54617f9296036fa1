"""Minimal PySCF stand-in (hydrogen-only molecules, STO-3G) for the reference's driver: see tests/shims/README.md."""
from . import dft, gto, scf  # noqa: F401

__version__ = "0.0-shim"
