"""B200-native forward/grounding path for T2S-QA (drop-in behind the pythia registry model API)."""
__version__ = "0.1.0"
