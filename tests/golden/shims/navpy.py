"""empty stub for navpy (image.py:9); geodesy is not on the matching path"""
