class Regularizer(object):
    def __call__(self, x):
        return 0.0


class L1L2(Regularizer):
    def __init__(self, l1=0.0, l2=0.0):
        self.l1, self.l2 = float(l1), float(l2)

    def get_config(self):
        return {"l1": self.l1, "l2": self.l2}


def l1(l=0.01):
    return L1L2(l1=l)


def l2(l=0.01):
    return L1L2(l2=l)


def get(identifier):
    return identifier


def serialize(r):
    from complexnn._layer import serialize_object
    return serialize_object(r)
