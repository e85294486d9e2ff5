"""Minimal stand-in for `getdist` (not installed, no network) used ONLY by
tests/golden/make_golden.py when it imports the reference from /root/reference.

loadMCSamples(root) concatenates every `<root>_<i>.txt` chain file in index
order with no burn-in removal (SURVEY.md F8b validated that this reproduces the
reference's own `initial_loglkl` fixture to 11 digits).
"""
import glob
import os
import re
import numpy as np

GELMAN_RUBIN_SCRIPT = []   # values popped by getGelmanRubin (harness-controlled)


class _Name:
    def __init__(self, name):
        self.name = name


class _ParamNames:
    def __init__(self, names):
        self.names = [_Name(n) for n in names]


class _Samples:
    def __init__(self, data, names):
        self.weights = data[:, 0]
        self.loglikes = data[:, 1]
        self.samples = data[:, 2:]
        self._names = names

    def getParamNames(self):
        return _ParamNames(self._names)

    def getGelmanRubin(self):
        if GELMAN_RUBIN_SCRIPT:
            return GELMAN_RUBIN_SCRIPT.pop(0)
        return 0.0


def loadMCSamples(root, **_):
    files = glob.glob(root + '_*.txt')
    def idx(f):
        m = re.search(r'_(\d+)\.txt$', f)
        return int(m.group(1)) if m else -1
    files = sorted([f for f in files if idx(f) >= 0], key=idx)
    data = np.concatenate([np.atleast_2d(np.loadtxt(f)) for f in files], axis=0)
    names = []
    with open(root + '.paramnames') as fh:
        for line in fh:
            line = line.strip()
            if line:
                names.append(line.split()[0])
    return _Samples(data, names)


class plots:  # pragma: no cover - plotting is out of scope
    @staticmethod
    def get_subplot_plotter(*a, **k):
        raise RuntimeError('plotting is out of scope')
