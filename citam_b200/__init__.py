"""citam_b200 — B200-native step loop for CITAM."""
