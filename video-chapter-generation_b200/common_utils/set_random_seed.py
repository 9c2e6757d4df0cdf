"""Mirror of the reference's common_utils/set_random_seed.py:6-10 (every script seeds with 123)."""
import random

import numpy as np
import torch


def use_fix_random_seed():
    np.random.seed(123)
    random.seed(123)
    torch.manual_seed(123)
    if torch.cuda.is_available():
        torch.cuda.manual_seed_all(123)
