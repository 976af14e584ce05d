"""Row f-1: the re-fitted schedule predictor keeps the reference's interface and feature
definition (sduss/worker/scheduler/policy/ESyMReD.py:20-53) and reproduces the B200 profile."""
import os

import numpy as np
import pytest

from sduss_b200.predictor import DATA_DIR, Predictor, default_model_path, features


def test_features_match_reference_definition():
    # ESyMReD.py:45-49: [n512, n768, n1024, 4 n512 + 9 n768 + 16 n1024, count of non-zero entries]
    f = features([[1, 0, 1], [0, 0, 0], [3, 2, 1]])
    assert f.tolist() == [[1, 0, 1, 20, 2], [0, 0, 0, 0, 0], [3, 2, 1, 46, 3]]


@pytest.mark.parametrize("model", ["sd3", "sdxl"])
def test_predictor_reproduces_profile(model):
    p = Predictor(default_model_path(model))
    d = np.loadtxt(os.path.join(DATA_DIR, f"unet_time_{model}_b200.csv"), delimiter=",", skiprows=1)
    pred = p.predict(d[:, :3].tolist()) * 50          # predict() is per step, the profile per 50 steps
    rel = np.abs(pred - d[:, 3]) / d[:, 3]
    # (the worst row is the single 512^2 SDXL request, 17 %: a latency-bound step the five linear features of
    #  the reference's predictor cannot bend to; the policy reads that case from STANDALONE, checked below)
    assert rel.mean() < 0.03 and rel.max() < 0.2, (rel.mean(), rel.max())
    # stand-alone latencies (Predictor.latency / esymred.json STANDALONE) come from the same profile
    for i, res in enumerate(("512", "768", "1024")):
        row = d[(d[:, :3] == np.eye(3)[i]).all(1)][0]
        assert p.get_latency(res) == pytest.approx(row[3] / 50)
        assert p.get_latency(int(res)) == p.get_latency(res)
    # more work never predicts less time
    base = p.predict([[2, 1, 1]])[0]
    for extra in ([3, 1, 1], [2, 2, 1], [2, 1, 2]):
        assert p.predict([extra])[0] > base


def test_predictor_rejects_unknown_model(tmp_path):
    import joblib
    path = tmp_path / "schedule_predictor_sd15.pkl"
    joblib.dump(object(), path)
    with pytest.raises(ValueError):   # same behaviour as the reference for an unknown model name
        Predictor(str(path))
