"""Drop-in `model` package: same import paths as the reference's model/ (model, model.gan, model.progressive_gan,
model.autoencoder, model.sdf_net).  Put this repo ahead of the reference on PYTHONPATH
(PYTHONPATH=/path/to/shapegan-b200:/path/to/shapegan) and the reference's train_*.py / demo_*.py import the
B200-native classes unchanged; util/datasets/rendering stay the reference's."""
from shapegan_b200.nn import (CHECKPOINT_PATH, LATENT_CODE_SIZE, LATENT_CODES_FILENAME, MODEL_PATH, BatchNorm1d, Lambda,  # noqa: F401
                              Linear, ReLU, SavableModule, Sequential, nn, os, torch)
