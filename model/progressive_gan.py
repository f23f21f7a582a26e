"""Drop-in for the reference's model/progressive_gan.py — re-exports the B200-native implementation."""
from shapegan_b200.nn import (CHECKPOINT_PATH, LATENT_CODE_SIZE, LATENT_CODES_FILENAME, MODEL_PATH, Lambda,  # noqa: F401
                              SavableModule)
from shapegan_b200.nn.progressive_gan import *  # noqa: F401,F403
from shapegan_b200.nn import progressive_gan as _impl

globals().update({k: v for k, v in vars(_impl).items() if not k.startswith('__')})
