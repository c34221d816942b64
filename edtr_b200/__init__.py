"""edtr_b200 — B200-native (sm_100a) implementation of EDTR's ControlLDM restore path."""
__version__ = "0.1.0"
