"""Synthetic inputs of the BASELINE configs: measurement sets following the
reference's python/generate_measurements.py recipe and the host-side assembly of the point
measurement matrix B (MeasuredOperator, linear_operator/measured_operator.cc:9-49,74-91)."""
import numpy as np


def sample_points(n, dim=2, dmin=0.1):
    """python/generate_measurements.py:51-66 (numpy default_rng(2154157), boundary distance > 0.1)."""
    rng = np.random.default_rng(seed=2154157)
    points = []
    while len(points) < n:
        x = np.asarray(rng.uniform(low=0.0, high=1.0, size=dim))
        if min(min(abs(x[d]), abs(1.0 - x[d])) for d in range(dim)) > 0.1 and all(np.linalg.norm(x - p) > dmin for p in points):
            points.append(x)
    return np.asarray(points)


def measurement_set(nmeas, dim=2, dmin=0.1):
    """(locations[nmeas, dim], sample_location[dim], mean[nmeas], variance[nmeas]) as printed by the
    reference script: means U[1,4] (seed 2513267), variances U[1,2] (seed 2511541)."""
    p = sample_points(nmeas + 1, dim, dmin)
    mean = np.random.default_rng(seed=2513267).uniform(size=nmeas, low=1.0, high=4.0)
    var = np.random.default_rng(seed=2511541).uniform(low=1.0, high=2.0, size=nmeas)
    return p[:-1], p[-1], mean, var


def nearest_vertex(nx, ny, x0):
    """measured_operator.cc:74-91: lexicographic index of the interior vertex closest to x0 (first
    minimum in lexicographic order)."""
    best = None
    ci, cj = int(np.floor(x0[0] * nx)), int(np.floor(x0[1] * ny))
    for j in sorted({min(max(cj + d, 1), ny - 1) for d in (0, 1)}):
        for i in sorted({min(max(ci + d, 1), nx - 1) for d in (0, 1)}):
            dist = np.sqrt((i * (1.0 / nx) - x0[0]) ** 2 + (j * (1.0 / ny) - x0[1]) ** 2)
            if best is None or dist < best[0]:
                best = (dist, (j - 1) * (nx - 1) + (i - 1))
    return best[1]


def point_measurement_matrix(nx, ny, locations, variance, variance_scaling=1.0):
    """COO triplets (rows, cols, vals, sigma) of B for point measurements (radius < 1e-12)."""
    rows = np.array([nearest_vertex(nx, ny, x0) for x0 in locations], dtype=np.int64)
    cols = np.arange(len(rows), dtype=np.int32)
    vals = np.ones(len(rows))
    return rows, cols, vals, variance_scaling * np.asarray(variance, dtype=np.float64)


def smooth_rhs(ctx_apply, nx, ny):
    """f = A u for the field u(x, y) = sin(pi x) sin(pi y)."""
    x = np.arange(1, nx) / nx
    y = np.arange(1, ny) / ny
    u = np.outer(np.sin(np.pi * y), np.sin(np.pi * x)).ravel()
    return ctx_apply(u)
