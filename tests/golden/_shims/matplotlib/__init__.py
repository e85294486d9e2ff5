"""No-op stand-in for matplotlib (not installed); plotting is out of scope."""
def use(*a, **k):
    pass
