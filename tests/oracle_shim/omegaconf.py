"""Stand-in for `omegaconf` built on PyYAML, with the float resolver OmegaConf has
(PyYAML alone reads `5e-4` as a string)."""
import re

import yaml


class _Loader(yaml.SafeLoader):
    pass


_Loader.add_implicit_resolver(
    "tag:yaml.org,2002:float",
    re.compile(r"^[-+]?(\d+\.?\d*|\.\d+)([eE][-+]?\d+)?$"),
    list("-+0123456789."),
)


class DictConfig(dict):
    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v


class OmegaConf:
    @staticmethod
    def load(path):
        with open(path) as f:
            return DictConfig(yaml.load(f, Loader=_Loader))

    @staticmethod
    def merge(*cfgs):
        out = DictConfig()
        for c in cfgs:
            out.update(c)
        return out

    @staticmethod
    def to_container(cfg, resolve=True):
        return dict(cfg)
