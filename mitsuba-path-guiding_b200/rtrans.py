"""Readers for the rough-transmittance tables (see tools/pack_rtrans.py for the packed layout)."""
import os
import struct

import numpy as np

from . import _abi as A

DATA_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data")


def load_packed(distribution):
    """Returns the table with ``raw`` re-interleaved in Mitsuba's file order
    (for i<2*etaN, j<alphaN: thetaN trans values then one diffuse value)."""
    name = "ggx" if distribution in (A.DISTR_GGX, "ggx") else "beckmann"
    buf = open(os.path.join(DATA_DIR, "rtrans_%s.bin" % name), "rb").read()
    assert buf[:8] == b"B200RTR1"
    etaN, alphaN, thetaN = struct.unpack_from("<iii", buf, 8)
    etaMin, etaMax, alphaMin, alphaMax = struct.unpack_from("<ffff", buf, 20)
    off = 36
    nt = 2 * etaN * alphaN * thetaN
    trans = np.frombuffer(buf, "<f4", nt, off).reshape(2 * etaN, alphaN, thetaN)
    diff = np.frombuffer(buf, "<f4", 2 * etaN * alphaN, off + 4 * nt).reshape(2 * etaN, alphaN)
    raw = np.concatenate([trans, diff[:, :, None]], axis=2).ravel().astype(np.float32)
    return dict(etaN=etaN, alphaN=alphaN, thetaN=thetaN, etaMin=etaMin, etaMax=etaMax, alphaMin=alphaMin,
                alphaMax=alphaMax, raw=raw, trans=trans, diff=diff)
