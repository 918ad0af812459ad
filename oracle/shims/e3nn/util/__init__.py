from . import _argtools, codegen, jit  # noqa: F401
