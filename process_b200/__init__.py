"""process_b200: B200-native read sampler behind ProCESS simulate_seq()/simulate_normal_seq()."""
from .forest import PhylogeneticForest  # noqa: F401
