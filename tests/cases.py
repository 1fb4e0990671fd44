"""Case tables shared by the CPU (oracle) and GPU (parity) tests -- must match oracle/make_golden.py."""
import importlib.util
import os

_spec = importlib.util.spec_from_file_location(
    "make_golden_cases", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_cases.py"))
_mod = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(_mod)
CONV_CASES, DENSE_CASES = _mod.CONV_CASES, _mod.DENSE_CASES


def conv_kwargs(rank, kw):
    """Normalise a case's kwargs to explicit values."""
    def tup(v):
        return (v,) * rank if isinstance(v, int) else tuple(v)
    return dict(strides=tup(kw.get("strides", 1)), padding=kw.get("padding", "valid"),
                data_format=kw.get("data_format", "channels_last") or "channels_last",
                dilation_rate=tup(kw.get("dilation_rate", 1)), activation=kw.get("activation"),
                use_bias=kw.get("use_bias", True))
