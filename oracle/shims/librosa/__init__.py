"""ORACLE SHIM: import-only stub of librosa (absent here; unused on the hot path)."""
