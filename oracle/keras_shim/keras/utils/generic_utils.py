def serialize_keras_object(instance):
    if instance is None:
        return None
    if hasattr(instance, "get_config"):
        return {"class_name": instance.__class__.__name__, "config": instance.get_config()}
    return getattr(instance, "__name__", str(instance))


def deserialize_keras_object(identifier, module_objects=None, custom_objects=None, printable_module_name="object"):
    raise NotImplementedError
