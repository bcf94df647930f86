"""gymnasium.spaces when available, otherwise minimal stand-ins with the attributes the
reference exposes (Discrete.n / sample / contains; Box.low / high / shape / dtype)."""

from __future__ import annotations

import numpy as np

try:  # pragma: no cover - gymnasium is not installed in the build image
    from gymnasium.spaces import Box, Discrete, Space  # type: ignore
except Exception:  # noqa: BLE001

    class Space:  # type: ignore[no-redef]
        pass

    class Discrete(Space):  # type: ignore[no-redef]
        def __init__(self, n, seed=None):
            self.n = int(n)
            self.shape = ()
            self.dtype = np.dtype(np.int64)
            self._rng = np.random.default_rng(seed)

        def sample(self):
            return int(self._rng.integers(0, self.n))

        def contains(self, x):
            try:
                return 0 <= int(x) < self.n
            except (TypeError, ValueError):
                return False

        __contains__ = contains

        def __repr__(self):
            return f"Discrete({self.n})"

        def __eq__(self, other):
            return isinstance(other, Discrete) and other.n == self.n

    class Box(Space):  # type: ignore[no-redef]
        def __init__(self, low, high, shape=None, dtype=np.float32):
            self.dtype = np.dtype(dtype)
            self.shape = tuple(shape) if shape is not None else np.asarray(low).shape
            self.low = np.broadcast_to(np.asarray(low, dtype=self.dtype), self.shape).copy()
            self.high = np.broadcast_to(np.asarray(high, dtype=self.dtype), self.shape).copy()

        def contains(self, x):
            x = np.asarray(x)
            return x.shape == self.shape and bool(np.all(x >= self.low) and np.all(x <= self.high))

        __contains__ = contains

        def __repr__(self):
            return f"Box({self.shape}, {self.dtype})"
