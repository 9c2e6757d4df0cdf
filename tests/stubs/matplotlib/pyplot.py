"""Test stub (see __init__.py)."""
