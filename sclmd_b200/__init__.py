"""sclmd_b200 -- B200-native (sm_100a) implementation of the sclmd generalized-Langevin /
NEGF hot path, behind the reference's Python class API (md, ebath, phbath, bpt, sig)."""
__version__ = "0.1.0"
