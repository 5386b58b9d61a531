"""Thin re-export shim: see compat/README.md."""
