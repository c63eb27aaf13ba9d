"""B200-native TTL (test-time low-rank adaptation of CLIP) hot path: Python host over libttl_b200.so."""
from .engine import ARCH_GEOMETRY, Engine, Hparams  # noqa: F401
