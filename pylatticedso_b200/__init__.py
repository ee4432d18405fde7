"""B200-native beam-FEM hot path for pyLatticeDSO lattices."""
__version__ = "0.1.0"
