"""Minimal stand-in for the `easydict` package (absent from this image) so the
reference's checkpoints (which pickle an EasyDict) can be loaded by
oracle/make_golden.py.  Test infrastructure only."""


class EasyDict(dict):
    def __getattr__(self, key):
        try:
            return self[key]
        except KeyError:
            raise AttributeError(key)

    def __setattr__(self, key, value):
        self[key] = value
