"""Minimal stand-in for the `timm` package: the reference's models/sr3_dwt.py:9 imports only
`timm.models.layers.DropPath`.  Used ONLY by tests/golden/make_golden.py to import the reference."""
