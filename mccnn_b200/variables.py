"""Named parameter store standing in for tf.get_variable in the builder (utils/MCConvBuilder.py:394-419) and the
dense glue (utils/MCNetworkUtils.py).

Initialisers follow the reference: weights = variance_scaling(factor=1.0, mode='FAN_AVG', uniform=True) -- i.e.
U(-l, l) with l = sqrt(3 / ((fan_in + fan_out)/2)), fan computed as TensorFlow does for rank >= 2 shapes
(receptive field = product of all but the last two dims) -- and zero biases.
"""
import math

import torch


def _fans(shape):
    if len(shape) < 1:
        return 1.0, 1.0
    if len(shape) == 1:
        return float(shape[0]), float(shape[0])
    if len(shape) == 2:
        return float(shape[0]), float(shape[1])
    rf = 1.0
    for d in shape[:-2]:
        rf *= d
    return float(shape[-2]) * rf, float(shape[-1]) * rf


class VariableStore:
    """name -> float32 tensor.  get_variable creates on first use and returns the stored tensor afterwards."""

    def __init__(self, device="cuda", seed=0):
        self.device = torch.device(device)
        self.vars = {}
        self.collections = {}
        self._gen = torch.Generator(device="cpu")
        self._gen.manual_seed(seed)

    def glorot_uniform(self, shape):
        fan_in, fan_out = _fans(shape)
        limit = math.sqrt(3.0 / max(1.0, (fan_in + fan_out) / 2.0))
        t = (torch.rand(tuple(shape), generator=self._gen, dtype=torch.float32) * 2.0 - 1.0) * limit
        return t.to(self.device)

    def zeros(self, shape):
        return torch.zeros(tuple(shape), dtype=torch.float32, device=self.device)

    def ones(self, shape):
        return torch.ones(tuple(shape), dtype=torch.float32, device=self.device)

    def get_variable(self, name, shape, initializer="glorot"):
        if name in self.vars:
            v = self.vars[name]
            if tuple(v.shape) != tuple(shape):
                raise ValueError("variable %s exists with shape %s, requested %s" % (name, tuple(v.shape), tuple(shape)))
            return v
        if initializer == "glorot":
            v = self.glorot_uniform(shape)
        elif initializer == "zeros":
            v = self.zeros(shape)
        elif initializer == "ones":
            v = self.ones(shape)
        else:
            raise ValueError("unknown initializer " + str(initializer))
        self.vars[name] = v
        return v

    def add_to_collection(self, name, value):
        """tf.add_to_collection; idempotent per tensor storage: the TensorFlow graph is built once, this eager code
        runs the builder on every forward, so re-adding the same variable (or a view of it) is a no-op."""
        coll = self.collections.setdefault(name, [])
        ptr = value.data_ptr() if hasattr(value, "data_ptr") else id(value)
        if all((c.data_ptr() if hasattr(c, "data_ptr") else id(c)) != ptr for c in coll):
            coll.append(value)

    def state_dict(self):
        return {k: v.detach().cpu() for k, v in self.vars.items()}

    def load_state_dict(self, sd):
        """Existing variables are overwritten IN PLACE (same storage, `_version` bumped): captured CUDA graphs, folded
        batch-norm caches keyed on the version counters, optimizer state and `requires_grad` all stay valid.  Unknown
        names are created."""
        with torch.no_grad():
            for k, v in sd.items():
                v = torch.as_tensor(v, dtype=torch.float32)
                cur = self.vars.get(k)
                if cur is None:
                    self.vars[k] = v.to(self.device).contiguous()
                elif tuple(cur.shape) != tuple(v.shape):
                    raise ValueError("variable %s has shape %s, the state dict holds %s" %
                                     (k, tuple(cur.shape), tuple(v.shape)))
                else:
                    cur.copy_(v)


_default = None


def default_store():
    global _default
    if _default is None:
        _default = VariableStore()
    return _default


def set_default_store(store):
    global _default
    _default = store
    return store
