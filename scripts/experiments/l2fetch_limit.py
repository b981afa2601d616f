import ctypes, os, sys, json
sys.path.insert(0, '/root/repo')
import torch
g = int(sys.argv[1])
torch.zeros(1, device='cuda')
rt = ctypes.CDLL('libcudart.so.12')
val = ctypes.c_size_t(0)
rt.cudaDeviceGetLimit(ctypes.byref(val), ctypes.c_int(5)); before = val.value
err = rt.cudaDeviceSetLimit(ctypes.c_int(5), ctypes.c_size_t(g))
rt.cudaDeviceGetLimit(ctypes.byref(val), ctypes.c_int(5))
print(f'limit before {before}, set {g}: err {err}, now {val.value}')
