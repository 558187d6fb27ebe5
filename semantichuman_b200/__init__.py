"""semantichuman_b200 -- B200-native (sm_100a) kernels and drop-in modules for SemanticHuman's spiral-mesh
autoencoder training step.  Importing this package loads libshb200.so and fails loudly if it is missing."""
from . import _capi  # noqa: F401  (loads the shared library; ImportError if it has not been built)
from .functions import PairLossLayout, l1_loss, pair_loss, partnorm_loss, pool, spiral_conv  # noqa: F401
from .indexing import PoolMatrix, SpiralGeometry  # noqa: F401
from .models import Pool, SpiralAutoencoder, SpiralAutoencoder_multiz_partkps, SpiralConv  # noqa: F401

__all__ = ["SpiralConv", "Pool", "SpiralAutoencoder", "SpiralAutoencoder_multiz_partkps", "SpiralGeometry",
           "PoolMatrix", "spiral_conv", "pool", "l1_loss", "partnorm_loss", "pair_loss", "PairLossLayout"]
