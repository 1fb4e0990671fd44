"""DECODA text-format loader restating working_example.py:19-66 (TEST INFRASTRUCTURE: fixture generation / tests).
Each line: 250 space-separated `r,i,j,k` tokens, TAB, 8 label tokens `v,v,v,v` (first value used)."""
import numpy as np


def load_decoda(filename, isquat=True):
    docs = open(filename, "r").readlines()
    x = np.zeros((len(docs), 250, 4 if isquat else 3))
    y = np.zeros((len(docs), 8))
    for d, doc in enumerate(docs):
        data, labels = doc.split("\t")[:2]
        for e, element in enumerate(data.split(" ")):
            comps = element.split(",")
            x[d, e] = [float(c) for c in (comps[:4] if isquat else comps[1:4])]
        for l, label in enumerate(labels.split(" ")):
            y[d, l] = float(label.split(",")[0])
    return x, y
