"""Tests-only stand-in for the `numpy-quaternion` package (requirements.txt:3 of the reference; not installed here, no
network).  The reference's optimisation scripts import it (`import quaternion as qu`, gradient-based-optimize.py:8) but the
code paths of the four paper tasks never call into it; the few constructors below exist so that an accidental use fails
loudly on anything beyond construction instead of silently."""
import numpy as np


class quaternion:  # noqa: N801 - the package's own spelling
    def __init__(self, w=1.0, x=0.0, y=0.0, z=0.0):
        self.w, self.x, self.y, self.z = float(w), float(x), float(y), float(z)

    def __repr__(self):
        return f"quaternion({self.w}, {self.x}, {self.y}, {self.z})"


def as_float_array(q):
    return np.array([q.w, q.x, q.y, q.z])


def from_float_array(a):
    a = np.asarray(a, dtype=np.float64)
    return quaternion(*a[:4])


def __getattr__(name):
    raise AttributeError(f"quaternion stand-in (tests/standins): '{name}' is not provided")
