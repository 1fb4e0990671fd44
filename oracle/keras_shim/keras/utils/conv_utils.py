def normalize_tuple(value, n, name):
    if isinstance(value, int):
        return (value,) * n
    value_tuple = tuple(value)
    if len(value_tuple) != n:
        raise ValueError("The `" + name + "` argument must be a tuple of " + str(n) + " integers. Received: " + str(value))
    for v in value_tuple:
        int(v)
    return value_tuple


def normalize_padding(value):
    padding = value.lower()
    if padding not in {"valid", "same", "causal"}:
        raise ValueError("The `padding` argument must be one of \"valid\", \"same\" (or \"causal\" for Conv1D). Received: "
                         + str(padding))
    return padding


def conv_output_length(input_length, filter_size, padding, stride, dilation=1):
    if input_length is None:
        return None
    assert padding in {"same", "valid", "full", "causal"}
    dilated = filter_size + (filter_size - 1) * (dilation - 1)
    if padding == "same":
        out = input_length
    elif padding == "valid":
        out = input_length - dilated + 1
    elif padding == "causal":
        out = input_length
    else:
        out = input_length + dilated - 1
    return (out + stride - 1) // stride
