"""Import-only stand-in for `import tensorflow as tf` (models/interspeech_model.py:30; never used there)."""
