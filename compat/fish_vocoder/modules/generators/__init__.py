"""Compat package: the reference dotted module paths, re-exporting vocoder_b200 classes (INTEGRATION.md option ii)."""
