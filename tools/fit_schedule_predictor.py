"""Fits the B200 schedule predictors from sduss_b200/data/unet_time_<model>_b200.csv (CPU only)
on the reference's five features (sduss_b200/predictor.py::features) and writes
  sduss_b200/data/schedule_predictor_<model>_b200.pkl   joblib, `.predict(X)` -> seconds / 50 steps
  sduss_b200/data/esymred_b200.json                     STANDALONE denoising table (esymred.json)
Candidates: the estimator the reference ships (sklearn MLPRegressor, hidden layers (32, 32, 16),
exp/schedule_predictor_*.pkl) and ordinary least squares; the one with the lower 5-fold
cross-validated relative error is kept (on B200 the step time is very nearly linear in the
features, and 70 compositions are few for the MLP).
Usage: python tools/fit_schedule_predictor.py"""
import json, os, sys, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import joblib
import numpy as np
from sklearn.linear_model import LinearRegression
from sklearn.neural_network import MLPRegressor
from sklearn.pipeline import make_pipeline
from sklearn.preprocessing import StandardScaler
from sduss_b200.predictor import DATA_DIR, features

class RelativeLeastSquares(LinearRegression):
    """Ordinary least squares on relative residuals (sample weight 1 / y^2): the policy compares
    predictions with deadlines that are multiples of the stand-alone latency, so a 10 % miss on a
    one-request batch matters as much as on a full one."""

    def fit(self, X, y):
        return super().fit(X, y, sample_weight=1.0 / np.square(y))


CANDIDATES = {
    "linear": lambda: LinearRegression(),
    "linear, relative residuals": lambda: RelativeLeastSquares(),
    "mlp(32,32,16)": lambda: make_pipeline(StandardScaler(), MLPRegressor(hidden_layer_sizes=(32, 32, 16),
                                                                         max_iter=5000, random_state=0)),
}


def cv_error(make, X, y):
    errs = []
    for f in range(5):
        te = np.arange(len(y)) % 5 == f
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            m = make().fit(X[~te], y[~te])
        errs += list(np.abs(m.predict(X[te]) - y[te]) / y[te])
    return float(np.mean(errs)), float(np.max(errs))


consts = {"STANDALONE": {}, "fit": {},
          "source": "tools/profile_compositions.py on 1x B200, this repo's denoising step, CFG on, seconds per 50 steps"}
for name in ("sd3", "sdxl"):
    d = np.loadtxt(os.path.join(DATA_DIR, f"unet_time_{name}_b200.csv"), delimiter=",", skiprows=1)
    X, y = features(d[:, :3]), d[:, 3]
    scores = {k: cv_error(mk, X, y) for k, mk in CANDIDATES.items()}
    pick = min(scores, key=lambda k: scores[k][0] + 0.25 * scores[k][1])
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        m = CANDIDATES[pick]().fit(X, y)
    if isinstance(m, RelativeLeastSquares):   # ship a plain sklearn estimator (no class of ours to unpickle)
        plain = LinearRegression()
        plain.coef_, plain.intercept_, plain.n_features_in_ = m.coef_, m.intercept_, m.n_features_in_
        m = plain
    joblib.dump(m, os.path.join(DATA_DIR, f"schedule_predictor_{name}_b200.pkl"))
    alone = {r: float(d[(d[:, :3] == np.eye(3)[i]).all(1), 3][0]) for i, r in enumerate(("512", "768", "1024"))}
    consts["STANDALONE"][name] = {"denoising": alone}
    consts["fit"][name] = {"compositions": int(len(y)), "estimator": pick,
                           "cv_relative_error_mean_max": {k: [round(v[0], 4), round(v[1], 4)] for k, v in scores.items()}}
    print(f"{name}: {len(y)} compositions; 5-fold CV relative error (mean, max): {scores}; kept {pick}; "
          f"standalone s/50 steps {alone}")
with open(os.path.join(DATA_DIR, "esymred_b200.json"), "w") as f:
    json.dump(consts, f, indent=1)
