"""Synthetic / example data of the BASELINE workloads (inputs only; no sampler logic).

logistic regression: X ~ N(0,1)/sqrt(P), beta* ~ N(0,1), y ~ Bernoulli(sigmoid(X beta*))
from numpy's Generator(PCG64(seed)).  Stock-Watson: the T = 252 observations of the reference's example
(WALNUTSpy_examples/StockWatson/swdata.json) stored as walnuts_b200/data/sw_y.npy (data, not code)."""
import os

import numpy as np


def synth_logreg(N=100_000, P=100, seed=0):
    rng = np.random.Generator(np.random.PCG64(seed))
    X = rng.standard_normal((N, P)) / np.sqrt(P)
    beta = rng.standard_normal(P)
    y = (rng.random(N) < 1.0 / (1.0 + np.exp(-(X @ beta)))).astype(np.float64)
    return X, y, beta


def stock_watson_series():
    return np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "sw_y.npy"))
