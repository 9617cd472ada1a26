"""A numpy-backed stand-in for the handful of torch calls bench.py's packed-image child makes, so that its logic can run
against the host-simulated engine on a box without a GPU (tests/test_engine_hostsim.py).  TEST INFRASTRUCTURE ONLY."""
import numpy as np

bfloat16 = "bfloat16"
uint8 = "uint8"


class Generator:
    def __init__(self, device=None):
        self.rng = np.random.default_rng(0)

    def manual_seed(self, seed):
        self.rng = np.random.default_rng(seed)
        return self


class _Tensor:
    def __init__(self, n, dtype):
        assert dtype == bfloat16
        self.a = np.zeros(n, dtype=np.uint16)

    def uniform_(self, lo, hi, generator=None):
        rng = generator.rng if generator else np.random.default_rng(0)
        f = rng.uniform(lo, hi, self.a.size).astype(np.float32)
        self.a[:] = (f.view(np.uint32) >> 16).astype(np.uint16)
        return self

    def copy_(self, other):
        self.a[:] = other.a
        return self

    def data_ptr(self):
        return self.a.ctypes.data


def empty(n, dtype=None, device=None, pin_memory=False):
    return _Tensor(n, dtype)


class cuda:
    @staticmethod
    def set_device(i):
        pass

    @staticmethod
    def synchronize():
        pass

    @staticmethod
    def device_count():
        import os

        return int(os.environ.get("HOSTSIM_DEVICES", "1"))
