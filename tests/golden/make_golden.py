#!/usr/bin/env python3
"""Convert the reference's committed data dumps into small text fixtures.

Run in the build container only (reads /root/reference, which does not exist on
the GPU box).  Outputs are plain CSV with full 17-digit precision so that the
`-m "not gpu"` and `-m gpu` suites can load them with numpy alone.

Sources (reference file -> fixture):
  Casadi/1exemplo.xlsx  (written by Casadi/multiple_shooting_casadi.py:325-334)
  Casadi/2exemplo.xlsx  (written by Casadi/single_shooting_v2.py:292-301)
  Casadi/3exemplo.xlsx  (written by mpctools/multiple_shooting_mpctools.py:141-150)
  Inverted_pendulum/invertpend_data_py.xlsx
                        (written by Inverted_pendulum/inverted_pendulum_single_shooting_mpctools.py:80-88)
  Trajectory Tracking/dados.csv, dados2.csv (Trajectory Tracking/Phiref.py:379-381)
  Trajectory Tracking/lane_change.csv       (input path of all lane-change scripts)
  Trajectory Tracking/out.csv               (extended path written by lane_change.py:72-79)

No openpyxl in the image: .xlsx is a zip of XML, parsed by hand.
"""
import os
import sys
import zipfile
import xml.etree.ElementTree as ET

import numpy as np

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))
NS = {"m": "http://schemas.openxmlformats.org/spreadsheetml/2006/main"}


def _col_index(ref):
    letters = "".join(ch for ch in ref if ch.isalpha())
    idx = 0
    for ch in letters:
        idx = idx * 26 + (ord(ch) - ord("A") + 1)
    return idx - 1


def read_xlsx(path):
    """Return (header list, float ndarray) of sheet1; first column (pandas index) dropped."""
    with zipfile.ZipFile(path) as z:
        shared = []
        if "xl/sharedStrings.xml" in z.namelist():
            root = ET.fromstring(z.read("xl/sharedStrings.xml"))
            for si in root.findall("m:si", NS):
                shared.append("".join(t.text or "" for t in si.iter("{%s}t" % NS["m"])))
        sheet = ET.fromstring(z.read("xl/worksheets/sheet1.xml"))
    rows = []
    for row in sheet.iter("{%s}row" % NS["m"]):
        cells = {}
        for c in row.findall("m:c", NS):
            v = c.find("m:v", NS)
            if c.get("t") == "inlineStr":
                val = "".join(t.text or "" for t in c.iter("{%s}t" % NS["m"]))
            elif v is None:
                continue
            else:
                val = v.text
                if c.get("t") == "s":
                    val = shared[int(val)]
            cells[_col_index(c.get("r"))] = val
        if cells:
            width = max(cells) + 1
            rows.append([cells.get(i, "") for i in range(width)])
    header = [h for h in rows[0][1:]]
    data = np.array([[float(x) for x in r[1:]] for r in rows[1:]], dtype=np.float64)
    return header, data


def save(name, header, data):
    path = os.path.join(OUT, name)
    np.savetxt(path, data, delimiter=",", header=",".join(header), comments="", fmt="%.17g")
    print("wrote", path, data.shape)


def main():
    if not os.path.isdir(REF):
        sys.exit("reference tree not present; fixtures are already committed")
    for src, dst in [
        ("Casadi/1exemplo.xlsx", "unicycle_ms_1exemplo.csv"),
        ("Casadi/2exemplo.xlsx", "unicycle_ss_2exemplo.csv"),
        ("Casadi/3exemplo.xlsx", "unicycle_mpctools_3exemplo.csv"),
        ("Inverted_pendulum/invertpend_data_py.xlsx", "pendulum_invertpend.csv"),
    ]:
        h, d = read_xlsx(os.path.join(REF, src))
        save(dst, h, d)
    for src, dst in [
        ("Trajectory Tracking/dados.csv", "lateral_ltv_dados.csv"),
        ("Trajectory Tracking/dados2.csv", "lateral_lti_dados2.csv"),
        ("Trajectory Tracking/lane_change.csv", "lane_change.csv"),
        ("Trajectory Tracking/out.csv", "lane_change_out.csv"),     # output of the path generator lane_change.py
    ]:
        with open(os.path.join(REF, src)) as f:
            header = f.readline().strip().split(",")
        d = np.loadtxt(os.path.join(REF, src), delimiter=",", skiprows=1, ndmin=2)
        if header[0] == "":  # pandas index column
            header, d = header[1:], d[:, 1:]
        save(dst, header, d)


if __name__ == "__main__":
    main()
