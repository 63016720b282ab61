"""Small graphs and exact Boltzmann laws shared by the move tests (CPU oracle and GPU)."""
import itertools

import numpy as np


def irregular_graph():
    """6 sites, degrees 1..4, mixed-sign real couplings, biases: nothing about it is symmetric."""
    edges = [((0, 1), -1.0), ((1, 2), 0.7), ((2, 0), 1.3), ((2, 3), -0.5), ((3, 4), 0.9), ((1, 4), -1.1),
             ((4, 5), 0.6)]
    biases = [0.2, -0.3, 0.0, 0.4, -0.1, 0.25]
    return edges, 6, biases


def boltzmann(edges, nvars, beta, biases=None):
    """-> (states bool[2^N, N] in the order of state_index, probabilities[2^N], energies[2^N])"""
    a = np.array([e[0][0] for e in edges]); b = np.array([e[0][1] for e in edges])
    j = np.array([e[1] for e in edges], dtype=float)
    bias = np.zeros(nvars) if biases is None else np.asarray(biases, float)
    bits = np.array(list(itertools.product([0, 1], repeat=nvars)), dtype=np.int64)[:, ::-1]  # site 0 = LSB
    s = 2.0 * bits - 1.0
    E = (s[:, a] * s[:, b] * j).sum(1) - s @ bias
    w = np.exp(-beta * (E - E.min()))
    return bits.astype(bool), w / w.sum(), E


def state_index(states):
    st = np.asarray(states, dtype=np.int64)
    return (st << np.arange(st.shape[-1])).sum(-1)


def histogram_z(states, probs):
    """z-score of every state's count against its exact probability (independent experiments)."""
    n = len(states)
    counts = np.bincount(state_index(states), minlength=len(probs))
    sigma = np.sqrt(n * probs * (1 - probs))
    return (counts - n * probs) / np.maximum(sigma, 1e-12)
