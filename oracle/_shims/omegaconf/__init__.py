"""Stand-in for `omegaconf` (not installed in this image), TEST INFRASTRUCTURE ONLY.

The reference (read-only under /root/reference) uses `DictConfig` as a type
annotation, for attribute get/set and for `dict(cfg)`; nothing else.  This shim is
only ever placed on sys.path by oracle/ref_runner.py when the live reference is
imported to generate golden fixtures.  It never ships on the product path.
"""


class DictConfig(dict):
    def __init__(self, d=None, **kw):
        super().__init__()
        for k, v in dict(d or {}, **kw).items():
            self[k] = DictConfig(v) if isinstance(v, dict) else v

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)

    def __setattr__(self, k, v):
        self[k] = v
