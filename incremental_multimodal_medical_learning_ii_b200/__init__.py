"""B200-native BioViL image-encoder + prompt-scorer hot path (sm_100a CUDA behind the reference's Python API)."""
__version__ = "0.1.0"
