"""B200-native GDRNPP per-ROI pose inference hot path."""
