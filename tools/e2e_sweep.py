"""Host-buffer path (qnn_conv_forward_host, cfg 2) timed for several pipeline chunk counts: python tools/e2e_sweep.py 1 2 4 8 16"""
import ctypes, os, sys, time
import numpy as np, torch
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(REPO, "quaternion-convolutional-neural-networks-for-end-to-end-automatic-speech-recognition_b200")
sys.path[:0] = [REPO, PKG]
from complexnn import _native
lib=_native.lib()
x=torch.randn(256,256,160).pin_memory(); y=torch.empty(256,256,256).pin_memory()
k=(np.random.default_rng(0).normal(size=(3,40,256))*0.05).astype(np.float32); b=np.zeros(256,np.float32)
desc=_native.make_conv_desc(1,256,(256,),40,64,(3,),(1,),(1,),"same","channels_last","relu")
hp=lambda a: ctypes.c_void_p(a.data_ptr() if hasattr(a,"data_ptr") else a.ctypes.data)
for ch in sys.argv[1:]:
    os.environ["QNN_HOST_CHUNKS"]=ch
    for _ in range(3): _native.check(lib.qnn_conv_forward_host(ctypes.byref(desc),hp(x),hp(k),hp(b),hp(y),None))
    t0=time.perf_counter()
    for _ in range(20): _native.check(lib.qnn_conv_forward_host(ctypes.byref(desc),hp(x),hp(k),hp(b),hp(y),None))
    print("chunks",ch,"%.3f ms"%((time.perf_counter()-t0)/20*1e3), flush=True)
