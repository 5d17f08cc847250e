"""Extract the reference's golden fields into small committed fixtures.

Run HERE (the container with /root/reference); the fixtures travel, the reference tree does not.
Source files (reference tree, test/references/):
  test_1p_cc-reference.vtu                    1p incompressible/compressible CCTpfa, 10x10   (field p)
  test_2p_incompressible_cc-reference.vtu     2p lens, 48x32, t = 3000 s                      (10 fields)
  test_2p_incompressible_tpfa_oilwet-reference.vtu  2p, oil-wet lens, no gravity, 9th output file    (10 fields)
  test_tracer_{explicit,implicit}_tpfa-reference.vtu  tracer/constvel, 50x50, t = 1e6 s (output 10), two components (D = 1e-8, 0)
  test_1ptracer_pressure-reference.vtu        1p on log-normal K, 50x50                       (p, permeability)
  test_1ptracer_transport-reference.vtu       tracer after 5000 s                             (x, X, rho, velocity)
  test_tracer_implicit_dispersion_tpfa-reference.vtu  tracer/constvel with Scheidegger dispersion (AlphaL 0.02, AlphaT 0.008), output 10
  test_1p_pointsources_timeindependent_cc-reference.vtu  1p with a 10 kg/s point source at the origin, 100x100 on [-1,1]^2  (p)
All are Float32 ASCII cell data in element (x-fastest) order; the reference's own comparison is
fuzzy: relative 1e-2, absolute 1.5e-7 (bin/testing/dumux_runtest.py / fuzzycomparevtu.py).
"""
import json
import os
import sys
import xml.etree.ElementTree as ET

import numpy as np

REF = os.environ.get("DUMUX_REFERENCE", "/root/reference")
OUT = os.path.dirname(os.path.abspath(__file__))

FILES = {
    "test_1p_cc": "test/references/test_1p_cc-reference.vtu",
    "test_2p_incompressible_cc": "test/references/test_2p_incompressible_cc-reference.vtu",
    "test_2p_incompressible_tpfa_oilwet": "test/references/test_2p_incompressible_tpfa_oilwet-reference.vtu",
    "test_tracer_explicit_tpfa": "test/references/test_tracer_explicit_tpfa-reference.vtu",
    "test_tracer_implicit_tpfa": "test/references/test_tracer_implicit_tpfa-reference.vtu",
    "test_1ptracer_pressure": "test/references/test_1ptracer_pressure-reference.vtu",
    "test_1ptracer_transport": "test/references/test_1ptracer_transport-reference.vtu",
    "test_1p_pointsources_timeindependent_cc": "test/references/test_1p_pointsources_timeindependent_cc-reference.vtu",
    "test_tracer_implicit_dispersion_tpfa": "test/references/test_tracer_implicit_dispersion_tpfa-reference.vtu",
}


def cell_data(path):
    root = ET.parse(path).getroot()
    piece = root.find("UnstructuredGrid/Piece")
    ncells = int(piece.get("NumberOfCells"))
    out = {}
    for da in piece.find("CellData"):
        name = da.get("Name")
        ncomp = int(da.get("NumberOfComponents", "1"))
        vals = np.array(da.text.split(), dtype=np.float32)
        assert vals.size == ncells * ncomp, (name, vals.size, ncells, ncomp)
        out[name] = vals.reshape(ncells, ncomp) if ncomp > 1 else vals
    return ncells, out


def main():
    manifest = {}
    for key, rel in FILES.items():
        path = os.path.join(REF, rel)
        ncells, data = cell_data(path)
        np.savez_compressed(os.path.join(OUT, key + ".npz"), **{k.replace(" ", "_").replace("^", "_").replace("/", "_").replace("(", "").replace(")", ""): v for k, v in data.items()})
        manifest[key] = {"source": rel, "cells": ncells, "fields": sorted(data.keys())}
    with open(os.path.join(OUT, "MANIFEST.json"), "w") as f:
        json.dump(manifest, f, indent=1, sort_keys=True)
    print(json.dumps(manifest, indent=1))


if __name__ == "__main__":
    sys.exit(main())
