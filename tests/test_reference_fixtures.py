"""Pins the oracle to numbers produced by the REFERENCE ITSELF (STOR-i/sgmcmc on R + TensorFlow), when they exist.

`tests/golden/make_reference_fixtures.R` generates `tests/golden/reference_v1.json` on a machine that has R, the reference package
and TensorFlow: estLogPost, per-tensor gradients, the full-data gradient at the mode and the control-variate estimate, on closed-form
inputs stored in the same file.  This image has neither R nor TensorFlow, the reference ships no golden vectors for the path
(SURVEY.md 8c), so until somebody runs that script the file is absent, this test is skipped and the oracle stays PARITY UNPINNED
(said so in DESIGN.md section 6 and in oracle/__init__.py)."""
import json
import os

import numpy as np
import pytest

FIXTURE = os.path.join(os.path.dirname(__file__), "golden", "reference_v1.json")
pytestmark = pytest.mark.skipif(not os.path.exists(FIXTURE), reason="reference_v1.json has not been generated (needs R + TensorFlow + the reference package)")


def _model(fam):
    from oracle import models as OM
    return {"gauss2d": OM.GaussMean(2, prior_sd=10.0), "mixture": OM.Mixture2(2, 10.0), "logistic5": OM.Logistic(5)}[fam]


@pytest.mark.parametrize("fam", ["gauss2d", "mixture", "logistic5"])
def test_oracle_matches_the_reference(fam):
    from oracle import samplers as OS
    from oracle import tfgraph as TG
    ref = json.load(open(FIXTURE))
    rec, idx = ref[fam], np.asarray(ref["meta"]["idx0"], dtype=np.int64)
    data = {k: np.asarray(v, dtype=np.float64) for k, v in rec["dataset"].items()}
    if "y" in data:
        data["y"] = data["y"].reshape(-1)
    params = {k: np.asarray(v, dtype=np.float64) for k, v in rec["params"].items()}
    theta2 = {k: np.asarray(v, dtype=np.float64).reshape(np.shape(params[k])) for k, v in rec["theta2"].items()}
    N, n = len(next(iter(data.values()))), int(rec["n"])
    om = _model(fam)
    batch = OS.take(data, idx)
    for dtype, rtol in ((np.float64, 2e-6), (np.float32, 1e-5)):
        post = OS.Posterior(om, N, n, dtype, use_prior=True)
        assert abs(post.est_logpost(params, batch) - rec["estLogPost"]) <= rtol * abs(rec["estLogPost"]) + 1e-5
        g = post.grad(params, batch)
        for k in params:
            want = np.asarray(rec["grad"][k], dtype=np.float64).reshape(np.shape(g[k]))
            assert np.abs(np.asarray(g[k], np.float64) - want).max() <= rtol * np.abs(want).max() + 1e-6, (fam, k, "gradient estimate")
        fg = post.full_grad(params, data)
        for k in params:
            want = np.asarray(rec["fullGrad"][k], dtype=np.float64).reshape(np.shape(fg[k]))
            assert np.abs(np.asarray(fg[k], np.float64) - want).max() <= 10 * rtol * max(np.abs(want).max(), 1.0), (fam, k, "full-data gradient")
    # control-variate estimate: the float32 op-by-op graph replica is the like-for-like comparison (fp32 cancellation)
    gp = TG.GraphPosterior(om, N, n, use_prior=True)
    full = gp.full_grad(params, data)
    at_hat, at_th = gp.grad(params, batch), gp.grad(theta2, batch)
    for k in params:
        want = np.asarray(rec["cvGrad"][k], dtype=np.float64).reshape(np.shape(at_th[k]))
        ours = np.asarray(full[k], np.float64) - np.asarray(at_hat[k], np.float64) + np.asarray(at_th[k], np.float64)
        scale = np.abs(np.asarray(at_th[k], np.float64)).max()
        assert np.abs(ours - want).max() <= 1e-5 * scale + 1e-6, (fam, k, "control-variate estimate")
