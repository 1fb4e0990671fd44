from . import conv_utils, generic_utils, np_utils, training_utils  # noqa: F401
