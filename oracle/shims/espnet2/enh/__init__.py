"""ORACLE SHIM package (test infrastructure)."""
