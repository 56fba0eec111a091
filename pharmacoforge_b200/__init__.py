"""pharmacoforge_b200: B200-native denoising hot path of PharmacoForge."""
__version__ = "0.1.0"
