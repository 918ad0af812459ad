"""``compile_mode`` is a TorchScript annotation only; a no-op for the eager oracle."""


def compile_mode(mode: str):
    def deco(cls):
        cls._e3nn_compile_mode = mode
        return cls

    return deco
