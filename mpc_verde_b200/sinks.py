"""Result sinks next to the hot path: the table layouts the reference scripts dump
(and their overlay scripts read back) and the tracking-error metrics of the lateral-error trackers, from the
arrays `closed_loop` returns.  `write_xlsx` writes the workbook `DataFrame.to_excel(path, sheet_name="Sheet1")`
produces (index column first, header row, numeric cells) with nothing but zipfile + XML — pandas/openpyxl are not
in this image — so that `Casadi/plot.py:6-8` (`pd.read_excel("1exemplo.xlsx")`, columns x, y, theta, v, w, t) and
`Inverted_pendulum/ploting.py` read GPU results like the scripts' own dumps; `write_csv` is the
`DataFrame.to_csv` twin (`Phiref.py:379-381`, read by `leitordados.py`).
"""
import zipfile
from xml.sax.saxutils import escape

import numpy as np


def unicycle_table(states, controls, T, n_steps):
    """Rows of `1exemplo.xlsx` / `2exemplo.xlsx` (Casadi/multiple_shooting_casadi.py:316-334): columns
    x, y, theta, v, w, t with q[0] = q[1] = x_init (the scripts stack the initial prediction first), the control
    column shifted one row up and its last row repeated, t = [0, 0, T, 2T, ...].
    states [n_steps+1, 3], controls [n_steps, 2] of ONE scenario as returned by closed_loop."""
    n = int(n_steps)
    q = np.vstack([states[0:1], states[: n]])                       # q[0] = q[1] = x_init, ..., state at iteration n-1
    w = np.vstack([controls[:n], controls[n - 1:n]])                # shifted: row i = control applied at iteration i
    t = np.concatenate([[0.0], np.arange(n) * T])
    return np.column_stack([q, w, t])


def write_csv(path, table, header):
    np.savetxt(path, table, delimiter=",", header=",".join(header), comments="", fmt="%.17g")


def _col_letters(i):
    s = ""
    i += 1
    while i:
        i, r = divmod(i - 1, 26)
        s = chr(ord("A") + r) + s
    return s


def write_xlsx(path, table, header, sheet_name="Sheet1", index=True):
    """The workbook of `pd.DataFrame(dict(zip(header, table.T))).to_excel(path, sheet_name=...)`
    (Casadi/multiple_shooting_casadi.py:325-334, single_shooting_v2.py:292-301,
    mpctools/multiple_shooting_mpctools.py:141-150, Inverted_pendulum/...:80-88): first column the 0-based row index
    under an empty header cell, then the named columns; numbers stored with 17 significant digits."""
    table = np.atleast_2d(np.asarray(table, dtype=np.float64))
    if table.shape[1] != len(header):
        raise ValueError("table has %d columns, header %d names" % (table.shape[1], len(header)))
    off = 1 if index else 0
    rows = []
    cells = "".join('<c r="%s1" t="inlineStr"><is><t>%s</t></is></c>' % (_col_letters(j + off), escape(str(h)))
                    for j, h in enumerate(header))
    rows.append('<row r="1">%s</row>' % cells)
    for i, r in enumerate(table):
        cells = '<c r="A%d"><v>%d</v></c>' % (i + 2, i) if index else ""
        cells += "".join('<c r="%s%d"><v>%s</v></c>' % (_col_letters(j + off), i + 2, repr(float(v))) for j, v in enumerate(r))
        rows.append('<row r="%d">%s</row>' % (i + 2, cells))
    ns = "http://schemas.openxmlformats.org/spreadsheetml/2006/main"
    rel = "http://schemas.openxmlformats.org/officeDocument/2006/relationships"
    pkg = "http://schemas.openxmlformats.org/package/2006"
    sheet = ('<?xml version="1.0" encoding="UTF-8" standalone="yes"?><worksheet xmlns="%s"><sheetData>%s</sheetData>'
             '</worksheet>' % (ns, "".join(rows)))
    workbook = ('<?xml version="1.0" encoding="UTF-8" standalone="yes"?><workbook xmlns="%s" xmlns:r="%s"><sheets>'
                '<sheet name="%s" sheetId="1" r:id="rId1"/></sheets></workbook>' % (ns, rel, escape(sheet_name)))
    content_types = ('<?xml version="1.0" encoding="UTF-8" standalone="yes"?><Types xmlns="%s/content-types">'
                     '<Default Extension="rels" ContentType="application/vnd.openxmlformats-package.relationships+xml"/>'
                     '<Default Extension="xml" ContentType="application/xml"/>'
                     '<Override PartName="/xl/workbook.xml" ContentType="application/vnd.openxmlformats-officedocument.'
                     'spreadsheetml.sheet.main+xml"/><Override PartName="/xl/worksheets/sheet1.xml" ContentType='
                     '"application/vnd.openxmlformats-officedocument.spreadsheetml.worksheet+xml"/></Types>' % pkg)
    root_rels = ('<?xml version="1.0" encoding="UTF-8" standalone="yes"?><Relationships xmlns="%s/relationships">'
                 '<Relationship Id="rId1" Type="%s/officeDocument" Target="xl/workbook.xml"/></Relationships>' % (pkg, rel))
    wb_rels = ('<?xml version="1.0" encoding="UTF-8" standalone="yes"?><Relationships xmlns="%s/relationships">'
               '<Relationship Id="rId1" Type="%s/worksheet" Target="worksheets/sheet1.xml"/></Relationships>' % (pkg, rel))
    with zipfile.ZipFile(path, "w", zipfile.ZIP_DEFLATED) as z:
        z.writestr("[Content_Types].xml", content_types)
        z.writestr("_rels/.rels", root_rels)
        z.writestr("xl/workbook.xml", workbook)
        z.writestr("xl/_rels/workbook.xml.rels", wb_rels)
        z.writestr("xl/worksheets/sheet1.xml", sheet)


UNICYCLE_COLUMNS = ("x", "y", "theta", "v", "w", "t")                       # 1/2/3exemplo.xlsx, Casadi/plot.py:9-25
PENDULUM_COLUMNS = ("x", "xdot", "theta", "thetadot", "u", "t")             # invertpend_data_py.xlsx
LATERAL_COLUMNS = ("x1", "x2", "x3", "u", "x", "y", "yref", "phiref", "rref", "deltaref")   # dados2.csv, Phiref.py:379


def lateral_table(states, controls, par, c, Delta):
    """Rows of `dados2.csv` (Phiref.py:360-381): plant state AFTER step t, the control of step t, the dead-reckoned path
    and the stage-0 references par[:, 0, t].  states [Nsim+1, 3], controls [Nsim], par [Nsim, Nt, 4] (device layout)."""
    n = controls.shape[0]
    xz = np.zeros(n)
    for t in range(1, n):
        xz[t] = xz[t - 1] + c[t] * np.cos(states[t, 1]) * Delta
    return np.column_stack([states[1:n + 1, :3], controls[:n], xz, states[:n, 0], par[:n, 0, :]])


def lateral_tracking_errors(x, u, par, a, b, c, Delta):
    """The error log of Trajectory Tracking/Trjectory_tracking_le_LTV.py:173-188, as written: dead-reckoned path
    (xz, yz), mean squared deviations from the stage-0 references and the running path distance `dist`,
    its running mean `mse` and maximum `max`.  x [Nsim+1, 3], u [Nsim], par [4, Nt, Nsim]."""
    Nsim = u.shape[0]
    xz, yz = [], []
    mean = np.zeros(4)
    dist = mse = mx = 0.0
    for t in range(Nsim):
        xz.append(0.0 if t == 0 else xz[t - 1] + c[t] * np.cos(x[t, 1]) * Delta)
        yz.append(x[t, 0])
        mean[0] += (x[t, 0] - par[0, 0, t]) ** 2 / Nsim
        mean[1] += (x[t, 1] - par[1, 0, t]) ** 2 / Nsim
        mean[2] += (x[t, 2] - par[2, 0, t]) ** 2 / Nsim
        mean[3] += (u[t] - par[3, 0, t]) ** 2 / Nsim
        dist += np.hypot(xz[t] - a[t], yz[t] - b[t]) / Nsim
        mse += dist / (t + 1)
        mx = max(mx, abs(dist))
    return {"path": np.array([xz, yz]), "mean_sq": mean, "dist": dist, "mse": mse, "max": mx}
