"""simvg_b200 — B200-native (sm_100a) implementation of SimVG's vision-language fusion train step.

Public surface = the reference's plugin API (`simvg_b200.models`: registries + BEIT3 / TextGuidedQuerySelectKDDETRHead /
MIXDETRMB) over hand-written CUDA reached through the C ABI in include/simvg_b200.h (libsimvg_b200.so)."""
__version__ = "0.1.0"
