"""mebt_b200 — B200-native (sm_100a) hot path of MeBT behind the reference's Python API."""
__all__ = ["ops"]
