#!/usr/bin/env python
"""Runs an UNMODIFIED script of the reference checkout (default: working_example.py) on top of this repository:
`complexnn` resolves to the B200 mirror package, `keras` / `tensorflow` to the facade.  Needs a CUDA device.

  python tools/run_reference_script.py /path/to/reference [script.py] [-- script args...]
e.g.  python tools/run_reference_script.py /root/reference working_example.py -- --model QDNN
"""
import os
import runpy
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(REPO, "quaternion-convolutional-neural-networks-for-end-to-end-automatic-speech-recognition_b200")


def main():
    if len(sys.argv) < 2:
        raise SystemExit(__doc__)
    ref = os.path.abspath(sys.argv[1])
    rest = sys.argv[2:]
    script = "working_example.py"
    if rest and rest[0] != "--":
        script, rest = rest[0], rest[1:]
    if rest and rest[0] == "--":
        rest = rest[1:]
    # our packages shadow the reference's `complexnn`; the reference's `models` package and data files stay visible
    sys.path[:0] = [PKG, os.path.join(PKG, "keras_facade")]
    sys.path.append(ref)
    import builtins
    if not hasattr(builtins, "xrange"):    # models/interspeech_model.py is Python-2 source (xrange(0, n/2)): a stand-in
        builtins.xrange = lambda *a: range(*[int(v) for v in a])   # lets getTimitModel2D run unchanged under Python 3
    os.chdir(ref)                      # the script opens 'decoda/...' relative to its checkout
    sys.argv = [script] + rest
    runpy.run_path(os.path.join(ref, script), run_name="__main__")


if __name__ == "__main__":
    main()
