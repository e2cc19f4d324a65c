"""Compat package (see compat/fish_vocoder/__init__.py)."""
