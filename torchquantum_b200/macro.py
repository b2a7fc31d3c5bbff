"""Constants of the reference's macro.py:31-50."""
import math
from string import ascii_lowercase

import numpy as np
import torch

C_DTYPE = torch.complex64
F_DTYPE = torch.float32
ABC = ascii_lowercase
ABC_ARRAY = np.array(list(ABC))
INV_SQRT2 = 1 / math.sqrt(2)
C_DTYPE_NUMPY = np.complex64
F_DTYPE_NUMPY = np.float32
