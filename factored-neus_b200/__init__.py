"""factored-neus_b200: B200-native (sm_100a) implementation of the Factored-NeuS
per-ray volume-rendering hot path behind the reference's NeuSRenderer / fields API
(reference: models/renderer.py, models/fields.py).  CUDA only -- there is no CPU fallback."""
from . import synthetic  # noqa: F401
from . import _lib  # noqa: F401
from .fields import (SDFNetwork, RenderingNetwork, SingleVarianceNetwork, RefColor, NeRF, Lvis,  # noqa: F401
                     IndirectLight)
from .renderer import NeuSRenderer  # noqa: F401
