"""chainer.link: ``Link`` / ``Chain`` with ``init_scope`` registration and ``namedparams`` in Chainer's path format
(``/predictor/layer_0/W``), which is what gives the reference's ``.npz`` files their key names
(``serializers.load_npz(file, L.Classifier(model))``, predict_folds.py:157,206)."""
import contextlib

import numpy as np

from .variable import Parameter


class Link:
    def __init__(self):
        self.__dict__["_params"] = []
        self.__dict__["_children"] = []
        self.__dict__["_within_init_scope"] = False
        self.name = None

    xp = np

    @contextlib.contextmanager
    def init_scope(self):
        old = self._within_init_scope
        self.__dict__["_within_init_scope"] = True
        try:
            yield
        finally:
            self.__dict__["_within_init_scope"] = old

    def __setattr__(self, name, value):
        if self.__dict__.get("_within_init_scope"):
            if isinstance(value, Parameter):
                value.name = name
                if name not in self._params:
                    self._params.append(name)
            elif isinstance(value, Link):
                value.name = name
                if name not in self._children:
                    self._children.append(name)
        object.__setattr__(self, name, value)

    def add_param(self, name, shape=None, initializer=None):
        with self.init_scope():
            setattr(self, name, Parameter(initializer, shape))

    def add_link(self, name, link):
        with self.init_scope():
            setattr(self, name, link)

    def params(self):
        for _, p in self.namedparams():
            yield p

    def namedparams(self, include_uninit=True):
        for name in sorted(self._params):
            p = self.__dict__[name]
            if include_uninit or p.data is not None:
                yield "/" + name, p
        for name in sorted(self._children):
            for path, p in self.__dict__[name].namedparams(include_uninit):
                yield "/" + name + path, p

    def children(self):
        for name in self._children:
            yield self.__dict__[name]

    def links(self):
        yield self
        for c in self.children():
            yield from c.links()

    def to_cpu(self):
        for c in self.children():
            c.to_cpu()
        return self

    def to_gpu(self, device=None):
        for c in self.children():
            c.to_gpu(device)
        return self

    def cleargrads(self):
        pass


class Chain(Link):
    def __init__(self, **links):
        super().__init__()
        for name, l in links.items():
            self.add_link(name, l)


class ChainList(Link):
    def __init__(self, *links):
        super().__init__()
        for i, l in enumerate(links):
            self.add_link(str(i), l)
