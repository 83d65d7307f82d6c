"""B200-native LightningDiT sampling engine (drop-in for the reference's model + transport sampler API)."""
from .lightningdit import LightningDiT, LightningDiT_models  # noqa: F401
from .transport import create_transport, Sampler, Transport, ModelType, PathType, WeightType  # noqa: F401
from .engine import Engine, load_library, LIB_PATH  # noqa: F401
from .checkpoint import load_checkpoint, load_weights_with_shape_check, select_weights, load_latent_stats, denormalize_latents  # noqa: F401
from .vavae import AutoencoderKL, VA_VAE, DecoderEngine  # noqa: F401
