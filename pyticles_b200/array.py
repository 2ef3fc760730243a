"""Particle storage: torch CUDA float64 tensors that accept numpy-style assignment.

pyticles stores particle state in numpy arrays and its scripts write into them with tuples
and arrays (`p.r[0,:] = (0.0, 0.0, 0.0)`, test/neighbour_list_test.py:17-19).  torch refuses
those right-hand sides, so the storage attributes are this thin Tensor subclass whose only
addition is a converting __setitem__.  Results of operations are plain tensors.
"""
import numpy as np
import torch


class PArray(torch.Tensor):
    __torch_function__ = torch._C._disabled_torch_function_impl

    def __setitem__(self, key, value):
        if not isinstance(value, torch.Tensor):
            if isinstance(value, (tuple, list, np.ndarray)):
                value = torch.as_tensor(np.asarray(value), dtype=self.dtype, device=self.device)
        elif value.device != self.device or value.dtype != self.dtype:
            value = value.to(device=self.device, dtype=self.dtype)
        if isinstance(key, np.ndarray):
            key = torch.as_tensor(key, device=self.device)
        torch.Tensor.__setitem__(self, key, value)

    def numpy(self):
        return self.detach().cpu().as_subclass(torch.Tensor).numpy()


def parray(t):
    return t.as_subclass(PArray)


def zeros(shape, device, dtype=torch.float64):
    return parray(torch.zeros(shape, dtype=dtype, device=device))


def ones(shape, device, dtype=torch.float64):
    return parray(torch.ones(shape, dtype=dtype, device=device))


def from_numpy(a, device, dtype=torch.float64):
    return parray(torch.as_tensor(np.ascontiguousarray(a), dtype=dtype).to(device))
