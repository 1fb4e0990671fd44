"""Import-only stand-ins (models/interspeech_model.py:10 imports them and never uses them)."""


class Callback(object):
    pass


class ModelCheckpoint(Callback):
    def __init__(self, *a, **k):
        pass


class LearningRateScheduler(Callback):
    def __init__(self, *a, **k):
        pass
