"""Accuracy metrics the reference appends to Experiment.txt (oracle = test infrastructure).

These are what ties the oracle to REFERENCE-HELD numbers: the `INITIAL MEASUREMENTS` block of the historic logs under
/root/reference/Data/Experiments/** is written right after triangulation (before any optimisation), so
"Av. error" / "RMSE" / "C1|C2 standard desv" there are functions of the simulated key points, unproject, the
triangulator and the metric only.  tests/golden/reference_pins.json holds a selection of them.

  measureSimAbsoluteMapErrors   Modules/Utils/Measurements.cc:8-98   (float accumulators, MapPoints 2j / 2j+1)
  calculatePixelsStandDev       Modules/Utils/Geometry.cc:370-498    (scenes.pixel_sigma)
"""
import numpy as np

from .f32 import f32, F


def sim_absolute_map_errors(X1, X2, original, moved):
    """Measurements.cc:8-98 -> (average movement, average error, RMSE) in millimetres.

    X1/X2: MapPoint positions (float32) of the j-th match in KF1/KF2 (MapPoints 2j and 2j+1 in insertion order,
    Mapping.cc:329-339); original/moved: ground-truth points of the same matches.  All sums are float32, in the
    reference's loop order."""
    X1, X2, original, moved = f32(X1), f32(X2), f32(original), f32(moved)
    n = X1.shape[0]
    total_movement = F(0)
    total_error = F(0)
    total_sq = F(0)

    def norm(v):
        return np.sqrt((v[0] * v[0] + v[1] * v[1]) + v[2] * v[2])

    def sqnorm(v):
        return (v[0] * v[0] + v[1] * v[1]) + v[2] * v[2]

    for j in range(n):
        movement = original[j] - moved[j]
        e1 = X1[j] - original[j]
        e2 = X2[j] - moved[j]
        total_movement = total_movement + norm(movement)
        total_error = total_error + (norm(e2) + norm(e1))
        total_sq = total_sq + (sqnorm(e1) + sqnorm(e2))
    count = 2 * n
    av_move = total_movement / F(n)
    av_err = total_error / F(count)
    rmse = np.sqrt(total_sq / F(count))
    return float(av_move * F(1000)), float(av_err * F(1000)), float(rmse * F(1000))
