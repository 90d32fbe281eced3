"""Inert stand-in for matplotlib (not installed in this image): the reference's training_util.py:6 imports pyplot at module
level and never calls it on the measured path.  Any attribute access on the stub raises, so a plotting call cannot pass silently."""
