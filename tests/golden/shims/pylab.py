"""empty stub for pylab (srtm.py:6 `from pylab import *`)"""
