class ImageDataGenerator(object):
    def __init__(self, *a, **k):
        raise NotImplementedError("image pipelines are outside the quaternion conv/dense path")
