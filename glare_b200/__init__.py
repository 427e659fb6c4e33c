"""glare_b200 -- B200 (sm_100a) kernels behind the reference GLARE operator interfaces.

Host code is Python/PyTorch (device memory, streams, torch.distributed); all hot-path arithmetic is in
``libglare_b200.so`` (``include/glare_b200.h``), reached through ctypes with raw device pointers.
There is no CPU or eager-PyTorch fallback for those operators: calling them without the library or on
CPU tensors raises.
"""
__version__ = "0.1.0"
