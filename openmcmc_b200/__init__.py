"""openmcmc_b200 — B200-native (sm_100a) engine for openMCMC's per-sweep hot path.

Mirrors the reference package layout (`mcmc`, `model`, `parameter`, `gmrf`, `distribution`, `sampler`) for the
samplers named in BASELINE.json; all numerics run in hand-written CUDA reached through libomc.so (include/omc.h).
"""

__version__ = "0.1.0"
