"""ORACLE SHIM."""
class InversibleInterface:
    pass
