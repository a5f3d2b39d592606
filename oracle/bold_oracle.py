"""CPU restatement of the post-integration stages: BOLD -> cut/filter/decimate -> FC -> GoF.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

  * bold_sim   : ``BOLDModel.Sim(rE, nnodes, dt)`` — external module, call site
                 /root/reference/netwWilsonCowanPlastic.py:144.  NOT in the
                 reference tree, not pinned, not installable: restated from the
                 published Balloon-Windkessel model (Friston et al. 2003 /
                 Stephan et al. 2007 parameterisation used by the same lab, see
                 SURVEY.md section 8c).  **parity unpinned** at bit level; pinned
                 statistically by output/*.txt.
  * sim_bold   : netwWilsonCowanPlastic.py:140-158 (cut Neq=2000, Bessel band-pass,
                 filtfilt, decimate).  Uses the same SciPy calls as the reference.
  * fc         : np.corrcoef(BOLD.T), whole_sweep_both.py:81
  * ssim       : skimage.metrics.structural_similarity defaults (call site
                 utils.py:48) restated from Wang et al. 2004 + scikit-image's
                 documented defaults (7x7 uniform window, sample covariance,
                 K1=0.01, K2=0.03, border of 3 cropped).  **parity unpinned**.
  * get_all_metrics : utils.py:42-50 (+ new_metric utils.py:28-31)
  * kuramoto   : utils.py:34-40
  * welch_peak : whole_sweep_both.py:90-95
"""
import numpy as np
from scipy import signal

# Balloon-Windkessel constants
BW = dict(kappa=1.0 / 0.65, gamma=1.0 / 0.41, tau=0.98, alpha=0.32, E0=0.4, V0=0.04, TE=0.04)
BW["k1"] = 4.3 * 40.3 * BW["E0"] * BW["TE"]
BW["k2"] = 25.0 * BW["E0"] * BW["TE"]
BW["k3"] = 1.0


def bold_sim(rE, nnodes=None, dt=0.04):
    """Balloon-Windkessel, explicit Euler, one step per row of rE.

    rE [T, N] (or [B, T, N]) -> BOLD same shape.  BOLD[i] is computed from the state
    BEFORE consuming rE[i].  State (s, f, v, q) starts at (0.1, 1, 1, 1).
    """
    rE = np.asarray(rE, dtype=np.float64)
    T = rE.shape[-2]
    shp = rE.shape[:-2] + rE.shape[-1:]
    s = np.full(shp, 0.1)
    f = np.ones(shp)
    v = np.ones(shp)
    q = np.ones(shp)
    out = np.empty_like(rE)
    ia = 1.0 / BW["alpha"]
    E0 = BW["E0"]
    for i in range(T):
        out[..., i, :] = BW["V0"] * (BW["k1"] * (1 - q) + BW["k2"] * (1 - q / v) + BW["k3"] * (1 - v))
        x = rE[..., i, :]
        va = v ** ia
        ds = x - BW["kappa"] * s - BW["gamma"] * (f - 1)
        df = s
        dv = (f - va) / BW["tau"]
        dq = (f * (1 - (1 - E0) ** (1 / f)) / E0 - q * va / v) / BW["tau"]
        s = s + dt * ds
        f = f + dt * df
        v = v + dt * dv
        q = q + dt * dq
    return out


def bessel_ba(bold_dt=0.04):
    """netwWilsonCowanPlastic.py:152 — returns (b, a) in SciPy's order (the reference names them a, b)."""
    return signal.bessel(2, [2 * 0.01 * bold_dt, 2 * 0.1 * bold_dt], btype="bandpass")


def filt_decimate(bold, BOLD_downsamp=1000, Neq=2000, bold_dt=0.04):
    """netwWilsonCowanPlastic.py:145-156 on an already computed BOLD [T, N]."""
    x = bold[Neq:, :]
    b, a = bessel_ba(bold_dt)
    y = signal.filtfilt(b, a, x, axis=0)
    return y[::BOLD_downsamp]


def sim_bold(E_t, nnodes=90, BOLD_downsamp=1000, dt=0.002, downsamp=20):
    """netwWilsonCowanPlastic.py:140-158."""
    return filt_decimate(bold_sim(E_t, nnodes, dt * downsamp), BOLD_downsamp, 2000, dt * downsamp)


def fc(bold):
    """whole_sweep_both.py:81."""
    return np.corrcoef(bold.T)


def flat_fc(FC):
    """utils.py:24-26 — row-major strict upper triangle."""
    n = len(FC)
    return np.concatenate([FC[i, i + 1:] for i in range(n)])


def _box7(x):
    """Mean over every 7x7 window fully inside the image -> [H-6, W-6]."""
    c = np.cumsum(np.cumsum(np.pad(x, ((1, 0), (1, 0))), axis=0), axis=1)
    return (c[7:, 7:] - c[:-7, 7:] - c[7:, :-7] + c[:-7, :-7]) / 49.0


def ssim(X, Y, data_range=1.0):
    """Mean SSIM, scikit-image defaults for float images (win 7, uniform, sample covariance)."""
    X = np.asarray(X, dtype=np.float64)
    Y = np.asarray(Y, dtype=np.float64)
    NP = 49.0
    cov_norm = NP / (NP - 1.0)
    ux, uy = _box7(X), _box7(Y)
    uxx, uyy, uxy = _box7(X * X), _box7(Y * Y), _box7(X * Y)
    vx = cov_norm * (uxx - ux * ux)
    vy = cov_norm * (uyy - uy * uy)
    vxy = cov_norm * (uxy - ux * uy)
    C1 = (0.01 * data_range) ** 2
    C2 = (0.03 * data_range) ** 2
    Smap = ((2 * ux * uy + C1) * (2 * vxy + C2)) / ((ux * ux + uy * uy + C1) * (vx + vy + C2))
    return float(Smap.mean())


def get_all_metrics(sFC, empFC, data_range=1.0):
    """utils.py:42-50 -> (corr, euc, ssim, new_metric)."""
    fe, fs = flat_fc(empFC), flat_fc(sFC)
    corr = np.corrcoef(fs, fe)[0, 1]
    euc = np.linalg.norm(fe - fs)
    new_metric = (1 - corr) + (fs.mean() - fe.mean()) ** 2      # utils.py:28-31
    return float(corr), float(euc), ssim(sFC, empFC, data_range), float(new_metric)


def kuramoto(sign):
    """utils.py:34-40."""
    ang = np.angle(signal.hilbert(sign, axis=0))
    k = np.abs(np.mean(np.exp(1j * ang), axis=1))
    return float(k.mean()), float(k.std())


def welch_peak(E_t, dt=0.002, nperseg=4000):
    """whole_sweep_both.py:90-95."""
    freqs, p = signal.welch(E_t.T, fs=1 / dt, nperseg=nperseg)
    m = p.mean(axis=0)
    return float(freqs[np.where(m == m.max())[0][0]])
