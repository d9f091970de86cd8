"""B200-native BEV input generation."""
