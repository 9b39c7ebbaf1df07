"""ORACLE SHIM: import-only stub of torch_complex (absent here)."""
