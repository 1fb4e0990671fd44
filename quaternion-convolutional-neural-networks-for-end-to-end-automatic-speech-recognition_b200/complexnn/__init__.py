"""`complexnn` -- drop-in mirror of the reference package of the same name (reference complexnn/__init__.py:9-17):
identical export list; the quaternion conv / dense layers run on hand-written sm_100a kernels via libqnn_b200.so."""
from .conv import (QuaternionConv,
                   QuaternionConv1D,
                   QuaternionConv2D,
                   QuaternionConv3D)
from .dense import QuaternionDense
from .init import (sqrt_init, qdense_init, qconv_init)
from .utils import (GetRFirst, GetIFirst, GetJFirst, GetKFirst, getpart_quaternion_output_shape_first,
                    get_rpart_first, get_ipart_first, get_jpart_first, get_kpart_first)
