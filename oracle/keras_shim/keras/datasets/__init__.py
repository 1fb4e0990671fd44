"""Import-only stand-ins (models/interspeech_model.py:11)."""
cifar10 = cifar100 = None
