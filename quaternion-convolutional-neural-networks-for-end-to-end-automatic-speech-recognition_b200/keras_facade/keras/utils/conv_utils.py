from complexnn._layer import normalize_tuple, normalize_padding, conv_output_length  # noqa: F401
