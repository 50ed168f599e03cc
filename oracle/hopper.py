"""Hopper-v5 oracle (oracle/hopper.c -> mjc_planar.h).  Oracle only; PARITY UNPINNED."""
from .mjc_planar import INFO_KEYS, OraclePlanar  # noqa: F401

NB, NQ, NV, NU, OBS = 5, 6, 6, 3, 11


class OracleHopper(OraclePlanar):
    robot = "hopper"
