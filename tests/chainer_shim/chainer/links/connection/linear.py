"""chainer/links/connection/linear.py: W (out, in) LeCunNormal by default, b (out,) zeros; ``in_size=None`` is learnt at
the first call (``in_size = x.size // x.shape[0]``); forward = F.linear = ``x.dot(W.T) + b``."""
from ...link import Link
from ...variable import Parameter


class Linear(Link):
    def __init__(self, in_size, out_size=None, nobias=False, initialW=None, initial_bias=None):
        super().__init__()
        from ...initializers import LeCunNormal
        if out_size is None:
            in_size, out_size = None, in_size
        self.out_size = out_size
        with self.init_scope():
            self.W = Parameter(LeCunNormal() if initialW is None else initialW)
            self.b = None if nobias else Parameter(0.0 if initial_bias is None else initial_bias, (out_size,))
        if in_size is not None:
            self.W.initialize((out_size, in_size))

    def __call__(self, x):
        from ... import functions as F
        if self.W.data is None:
            self.W.initialize((self.out_size, x.size // x.shape[0]))
        return F.linear(x, self.W, self.b)
