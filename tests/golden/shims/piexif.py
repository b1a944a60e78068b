"""empty stub for piexif (exif.py:6)"""
