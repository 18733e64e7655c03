"""B200-native spectral front-end for Audio-Style-Transfer (STFT + CQT + normalise, iSTFT, stats)."""
__version__ = "0.1.0"
