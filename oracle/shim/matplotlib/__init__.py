"""Stub: humanoid/utils/logger.py imports matplotlib.pyplot at module import."""
