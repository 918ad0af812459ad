"""``CodeGenMixin`` only manages pickling of generated fx graphs in e3nn; the eager oracle needs none of it."""


class CodeGenMixin:
    pass
