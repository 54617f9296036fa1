"""scf.diis.CDIIS as dft.py:184,225 uses it: Pulay DIIS on the commutator error S D F - F D S."""
from scf_driver import CDIIS  # noqa: F401  (tests/scf_driver.py)
