"""empty stub: the reference imports matplotlib but never calls it on the matching path"""
