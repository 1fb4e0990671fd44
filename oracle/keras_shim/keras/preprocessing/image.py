"""Import-only stand-in (models/interspeech_model.py:20)."""


class ImageDataGenerator(object):
    def __init__(self, *a, **k):
        raise NotImplementedError("outside the quaternion conv/dense path")
