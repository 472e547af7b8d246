"""seg2eye_b200 -- B200-native (sm_100a) drop-in for the SPADE+Style G/D training step of mcbuehler/Seg2Eye.

    import seg2eye_b200; seg2eye_b200.install_dropin()

makes the reference's `import models`, `import models.networks`, `from models.pix2pix_model import Pix2PixModel`
and `from trainers.pix2pix_trainer import Pix2PixTrainer` resolve to this package."""
import importlib
import sys

__version__ = "0.1.0"


def install_dropin():
    """Alias the reference's import paths for the hot path to this package (train.py:14-20, test.py:8-11)."""
    names = {
        "models": "seg2eye_b200.models",
        "models.networks": "seg2eye_b200.models.networks",
        "models.pix2pix_model": "seg2eye_b200.models.pix2pix_model",
        "trainers": "seg2eye_b200.trainers",
        "trainers.pix2pix_trainer": "seg2eye_b200.trainers.pix2pix_trainer",
    }
    for alias, target in names.items():
        sys.modules[alias] = importlib.import_module(target)
    for sub in ("base_network", "generator", "discriminator", "encoder", "normalization", "architecture", "loss"):
        sys.modules["models.networks." + sub] = importlib.import_module("seg2eye_b200.models.networks." + sub)
