def get(identifier):
    return identifier


def serialize(constraint):
    if constraint is None:
        return None
    return {"class_name": constraint.__class__.__name__, "config": {}}
