"""Shim (import-time only)."""
