"""vcg_b200: Python host side of the B200-native chapter-boundary scorer (ctypes over libvcg_b200.so)."""
from . import binding  # noqa: F401
