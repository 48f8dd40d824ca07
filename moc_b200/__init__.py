"""moc_b200: B200-native (sm_100a) implementation of the MOC per-slide hot path."""
__version__ = "0.1.0"
