"""Stub of svgwrite (TEST INFRASTRUCTURE ONLY): only the Drawing name is imported."""
