"""Walker2d-v5 oracle (oracle/walker2d.c -> mjc_planar.h).  Oracle only; PARITY UNPINNED."""
from .mjc_planar import INFO_KEYS, OraclePlanar  # noqa: F401

NB, NQ, NV, NU, OBS = 8, 9, 9, 6, 17


class OracleWalker2d(OraclePlanar):
    robot = "walker2d"
