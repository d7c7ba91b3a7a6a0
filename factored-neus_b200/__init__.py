"""factored-neus_b200: B200-native (sm_100a) implementation of the Factored-NeuS
per-ray volume-rendering hot path behind the reference's NeuSRenderer / fields API."""
from . import synthetic  # noqa: F401
