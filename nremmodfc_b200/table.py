"""The reference's on-disk result format (whole_sweep_both.py:112-116) and its consumers' reduction.

* ``write_rank_file``  one tab-separated row per simulation with the reference's 20 columns and ``:.4f`` formatting,
                       header on first write — what each SLURM rank appends to ``output/temp/..._rank{r}``.
* ``collapse``         the hand "collapse" step (whole_sweep_both.py:32): concatenate rank files into the
                       comma-separated table that heatmaps.py:31 / figures read with ``pd.read_csv``.
* ``euccorr_optima``   heatmaps.py:28-58: cell means of ``e/|corr|`` inside the window
                       ``thx=(-0.08, 0.2), thy=(-0.2, 0.08)`` and the arg-min per state.
"""
import os

import numpy as np

COLUMNS = ["rank", "seed", "delta_G", "delta_sigma", "ssimW", "ssimN1", "ssimN2", "ssimN3", "corrW", "corrN1", "corrN2",
           "corrN3", "eW", "eN1", "eN2", "eN3", "sync", "meta", "mean", "peakfreq"]
STATES = ("W", "N1", "N2", "N3")


def rows_from_sweep(out, rank, seeds, delta_G, delta_sigma, peakfreq=None):
    """Assemble [B, 20] rows from ``SweepPlan.run`` output (gof[b, state, (corr, euc, ssim, new)])."""
    gof = np.asarray(out["gof"])
    B = gof.shape[0]
    pf = np.full(B, np.nan) if peakfreq is None else np.asarray(peakfreq, dtype=np.float64)
    return np.column_stack([np.full(B, rank), np.asarray(seeds), np.asarray(delta_G), np.asarray(delta_sigma),
                            gof[:, :, 2], gof[:, :, 0], gof[:, :, 1], out["sync"], out["meta"], out["mean"], pf])


def format_row(row):
    """One line exactly as whole_sweep_both.py:116 writes it (rank and seed as integers, 18 floats with 4 decimals)."""
    return "\t".join([str(int(row[0])), str(int(row[1]))] + [f"{v:.4f}" for v in row[2:]]) + "\n"


def write_rank_file(path, rows):
    """Append rows to a rank file, writing the header if the file does not exist (whole_sweep_both.py:112-116)."""
    new = not os.path.isfile(path)
    with open(path, "a") as f:
        if new:
            f.write("\t".join(COLUMNS) + "\n")
        for r in np.atleast_2d(rows):
            f.write(format_row(r))


def collapse(rank_files, out_path):
    """Concatenate tab-separated rank files into the comma-separated table the figures read."""
    with open(out_path, "w") as out:
        out.write(",".join(COLUMNS) + "\n")
        for p in rank_files:
            with open(p) as f:
                for i, line in enumerate(f):
                    if i == 0 or not line.strip():
                        continue
                    out.write(",".join(line.rstrip("\n").split("\t")) + "\n")


def read_table(path):
    """Comma-separated collapsed table -> dict of column arrays."""
    a = np.genfromtxt(path, delimiter=",", names=True)
    return {n: a[n] for n in a.dtype.names}


def euccorr_optima(tab, thx=(-0.08, 0.2), thy=(-0.2, 0.08)):
    """heatmaps.py:28-58: per state, the (delta_G, delta_sigma) cell minimising the cell mean of e/|corr| inside
    the window; returns {state: (delta_G, delta_sigma, value)}."""
    dG, dS = np.round(tab["delta_G"], 4), np.round(tab["delta_sigma"], 4)
    res = {}
    for s in STATES:
        ec = tab[f"e{s}"] / np.abs(tab[f"corr{s}"])
        best = None
        for g in np.unique(dG):
            for sg in np.unique(dS):
                if not (thx[0] <= g <= thx[1] and thy[0] <= sg <= thy[1]):
                    continue
                m = ec[(dG == g) & (dS == sg)]
                if len(m) and (best is None or m.mean() < best[2]):
                    best = (float(g), float(sg), float(m.mean()))
        res[s] = best
    return res
