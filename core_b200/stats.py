"""ma::stats-compatible tables from the arrays a sweep leaves on the host.

Reference: ma/maStats.cc:12-45 (what goes into the vectors: owned edges' metric lengths; cbrt -- 2-D: signed sqrt -- of
the owned simplex elements' mean-ratio qualities, both in mesh iteration order) and test/measureAnisoStats.cc:108-122,
217-243 (the files: <prefix>/linear_tables/linearETable_<rank>.dat and linearQTable_<rank>.dat, one value per line
through `std::ostream << double`, i.e. printf("%g")).
"""
import ctypes as C
import os

import numpy as np

from ._lib import lib


def linear_qualities(qualities, elem_owned=None, simplex=None, dim=3):
    """ma::getLinearQualitiesInMetricSpace (maStats.cc:12-31) from the per-element qualities of a sweep."""
    q = np.ascontiguousarray(qualities, dtype=np.float64)
    keep = None
    if elem_owned is not None or simplex is not None:
        keep = np.ones(len(q), bool)
        if elem_owned is not None:
            keep &= np.asarray(elem_owned).astype(bool)
        if simplex is not None:
            keep &= np.asarray(simplex).astype(bool)
        keep = np.ascontiguousarray(keep, dtype=np.uint8)
    out = np.empty(len(q))
    n = C.c_int64(0)
    # the host's libm cbrt, as the reference calls it (numpy's own cbrt differs from glibc's in the last bit)
    rc = lib().mag_linear_qualities(int(dim), len(q), q.ctypes.data, None if keep is None else keep.ctypes.data,
                                    out.ctypes.data, C.byref(n))
    if rc:
        raise ValueError("mag_linear_qualities: bad arguments")
    return out[:n.value].copy()


def edge_lengths(lengths, edge_owned=None):
    """ma::getEdgeLengthsInMetricSpace (maStats.cc:33-45): owned edges, iteration order."""
    L = np.asarray(lengths, dtype=np.float64)
    return L if edge_owned is None else L[np.asarray(edge_owned).astype(bool)]


def write_table(path, values):
    """writeTable of test/measureAnisoStats.cc:108-122 for a one-column table."""
    with open(path, "w") as f:
        f.write("".join("%g\n" % v for v in values))


def write_linear_tables(prefix, rank, lengths, qualities, edge_owned=None, elem_owned=None, simplex=None, dim=3):
    """<prefix>/linear_tables/linear{E,Q}Table_<rank>.dat as measureAnisoStats writes them; returns the two paths."""
    d = os.path.join(prefix, "linear_tables")
    os.makedirs(d, exist_ok=True)
    pe, pq = os.path.join(d, "linearETable_%d.dat" % rank), os.path.join(d, "linearQTable_%d.dat" % rank)
    write_table(pe, edge_lengths(lengths, edge_owned))
    write_table(pq, linear_qualities(qualities, elem_owned, simplex, dim))
    return pe, pq
