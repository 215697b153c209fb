"""ibc_b200: B200-native (sm_100a) implementation of the energy-based-policy hot
path of google-research/ibc.  `ibc_b200.ibc.*` and `ibc_b200.networks.*` mirror
the reference module paths; `ibc_b200.engine` is the torch-facing wrapper of
the C-ABI (include/ibc_b200.h)."""
__version__ = '0.1.0'
