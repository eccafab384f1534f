"""mm-pihm_b200: B200-native MM-PIHM hot path (CVODE RHS + implicit-integrator
vector work).  Host-side Python mirror of the C ABI in include/pihm_b200.h."""
from . import watershed  # noqa: F401
