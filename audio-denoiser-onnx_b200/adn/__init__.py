"""adn: host side of the B200-native speech-denoising path."""
from . import _lib, chunker, export, gtcrn_params, metadata, modelfile, stft_tables  # noqa: F401
from .model import Model  # noqa: F401

__all__ = ["Model", "export", "chunker", "metadata", "modelfile", "stft_tables", "gtcrn_params"]
