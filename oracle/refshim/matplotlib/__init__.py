"""stub: the reference imports matplotlib.mlab / pyplot at module scope only."""
mlab = None
