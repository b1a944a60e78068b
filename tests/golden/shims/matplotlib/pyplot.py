"""empty stub (matcher.py:8 imports pyplot, never used on the hot path)"""
