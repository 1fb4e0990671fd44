from complexnn._layer import serialize_object as serialize_keras_object  # noqa: F401


def deserialize_keras_object(identifier, module_objects=None, custom_objects=None, printable_module_name="object"):
    raise NotImplementedError
