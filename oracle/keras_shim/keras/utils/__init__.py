from . import conv_utils, generic_utils  # noqa
