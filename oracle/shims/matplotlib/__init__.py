"""Stub of matplotlib (TEST INFRASTRUCTURE ONLY): only cm.get_cmap / colors.to_hex names."""
