"""ORACLE SHIM: import-only stub (the hot path never instantiates ComplexTensor)."""
class ComplexTensor:  # pragma: no cover
    def __init__(self, real, imag):
        self.real, self.imag = real, imag
