import sys, os, ctypes, importlib
import numpy as np
sys.path.insert(0, os.getcwd())
import torch
from elasticdeform_b200 import _lib
dg = importlib.import_module("elasticdeform_b200.deform_grid")
lib = _lib.load_library(); dev = torch.device("cuda", 0)
X = torch.rand((256,)*3, device=dev); Y = torch.empty_like(X)
for ax in (0, 2, 0, 2):
    dg._spline_filter1d_device(lib, X, Y, ax, 3)
torch.cuda.synchronize(); print("done")
