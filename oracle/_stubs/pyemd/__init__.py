"""Empty stand-in for `pyemd` (absent here): utils/loader.py imports
evaluation.mmd, which imports pyemd at module scope.  Never called."""


def emd(*args, **kwargs):
    raise NotImplementedError("pyemd is not available in this image")
