"""B200-native quaternion convolution / dense hot path.

Directory layout (the directory name is not an importable identifier; put this directory on sys.path -- tests/conftest.py,
bench.py and __graft_entry__.py do -- and `import complexnn`, exactly as users of the reference do):

  csrc/        hand-written sm_100a CUDA (tcgen05 / TMEM / TMA) + the C ABI declared in include/qnn.h
  lib/         libqnn_b200.so, built in-tree by build.py (git-ignored)
  complexnn/   host-side mirror of the reference's `complexnn` package: same names, same constructor arguments,
               same stored-weight layout, `call` goes through ctypes into the library
"""
