"""pastix_b200 — B200-native sopalin numeric phase (factorization + up_down) for PaStiX 5.2."""
from .sopalin import Sopalin, SolverMatrix, PastixB200Error, critere_from_norm, FACTO, FLTTYPE, DTYPE  # noqa: F401
