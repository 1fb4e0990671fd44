"""Import-only stand-in so that `import tensorflow as tf` in models/interspeech_model.py succeeds; nothing in the
quaternion conv/dense path uses TensorFlow."""
__version__ = "0.0-b200-facade"
