"""Schedule predictor for sduss' `esymred` policy, re-fitted for the B200 denoising step
(SURVEY.md section 8f, row f-1).

Drop-in for `Predictor` in sduss/worker/scheduler/policy/ESyMReD.py:20-53: same constructor
argument (path of a joblib pickle whose name contains "sdxl" or "sd3"), same `predict` features
([n512, n768, n1024, 4 n512 + 9 n768 + 16 n1024, number of non-zero counts]) and output (seconds
per denoising step: the model is trained on seconds per 50 steps and divided by 50), same
`get_latency(resolution)` (stand-alone seconds per step). The reference's pickles and constants
were measured on H100 with its own kernels; on B200 they over-estimate the step by 2.3-3.1x, which
makes the policy abort requests that would have met their deadline. The data here comes from
tools/profile_compositions.py (this repo's step, CFG on), the fit from
tools/fit_schedule_predictor.py. Host-side scheduling code: no GPU work happens here."""
import json
import os
from typing import List

import numpy as np

DATA_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data")
STEPS = 50  # the profiles (and the reference's) are seconds per 50 denoising steps


def features(task_distribute) -> np.ndarray:
    """ESyMReD.py:45-49: counts, patch-weighted load (4/9/16 patches of 256 px per image), and the
    number of distinct resolutions in the batch."""
    t = np.asarray(task_distribute, dtype=np.float64).reshape(-1, 3)
    load = t[:, :1] * 4 + t[:, 1:2] * 9 + t[:, 2:3] * 16
    nz = np.count_nonzero(t, axis=1).reshape(-1, 1)
    return np.concatenate((t, load, nz), axis=1)


def default_model_path(model_name: str) -> str:
    return os.path.join(DATA_DIR, f"schedule_predictor_{model_name}_b200.pkl")


class Predictor:
    def __init__(self, model_path: str):
        import joblib
        self.model = joblib.load(model_path)
        self.model_path = model_path
        name = os.path.basename(model_path)
        with open(os.path.join(DATA_DIR, "esymred_b200.json")) as f:
            consts = json.load(f)
        if "sdxl" in name:
            key = "sdxl"
        elif "sd3" in name:
            key = "sd3"
        else:
            raise ValueError()
        self.latency = {res: sec / STEPS for res, sec in consts["STANDALONE"][key]["denoising"].items()}

    def get_latency(self, resolution):
        return self.latency[str(resolution)]

    def predict(self, task_distribute: List):
        return self.model.predict(features(task_distribute)) / STEPS
