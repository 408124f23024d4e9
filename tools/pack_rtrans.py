#!/usr/bin/env python
"""Repack Mitsuba's precomputed rough-transmittance tables (data/microfacet/{beckmann,ggx}.dat,
format documented in src/bsdfs/rtrans.h:81-150) into this repo's own de-interleaved layout:

    magic "B200RTR1" | i32 etaN, alphaN, thetaN | f32 etaMin, etaMax, alphaMin, alphaMax |
    f32 trans[2*etaN][alphaN][thetaN] | f32 diff[2*etaN][alphaN]

These are DATA tables required by the `roughplastic` model definition (roughplastic.cpp:290-307),
not source code. Usage: python tools/pack_rtrans.py /path/to/mitsuba/data/microfacet <outdir>
"""
import struct
import sys

import numpy as np


def load_dat(path):
    raw = open(path, "rb").read()
    hdr = b"MTS_TRANSMITTANCE"
    assert raw[: len(hdr)] == hdr, "bad header"
    off = len(hdr)
    etaN, alphaN, thetaN = struct.unpack_from("<QQQ", raw, off)
    off += 24
    etaMin, etaMax, alphaMin, alphaMax = struct.unpack_from("<ffff", raw, off)
    off += 16
    n = 2 * etaN * alphaN * (thetaN + 1)
    data = np.frombuffer(raw, dtype="<f4", count=n, offset=off)
    assert off + 4 * n == len(raw)
    return dict(etaN=etaN, alphaN=alphaN, thetaN=thetaN, etaMin=etaMin, etaMax=etaMax, alphaMin=alphaMin,
                alphaMax=alphaMax, raw=data)


def main():
    src, dst = sys.argv[1], sys.argv[2]
    for name in ("beckmann", "ggx"):
        t = load_dat("%s/%s.dat" % (src, name))
        a = t["raw"].reshape(2 * t["etaN"], t["alphaN"], t["thetaN"] + 1)
        trans = np.ascontiguousarray(a[:, :, : t["thetaN"]])
        diff = np.ascontiguousarray(a[:, :, t["thetaN"]])
        with open("%s/rtrans_%s.bin" % (dst, name), "wb") as f:
            f.write(b"B200RTR1")
            f.write(struct.pack("<iii", t["etaN"], t["alphaN"], t["thetaN"]))
            f.write(struct.pack("<ffff", t["etaMin"], t["etaMax"], t["alphaMin"], t["alphaMax"]))
            f.write(trans.astype("<f4").tobytes())
            f.write(diff.astype("<f4").tobytes())
        print(name, trans.shape, diff.shape)


if __name__ == "__main__":
    main()
