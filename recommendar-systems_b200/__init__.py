"""B200-native graph-propagation hot path for MMRec-style recommenders."""
__version__ = "0.1.0"
