"""Constant-velocity Kalman filter over (cx, cy, aspect, h, and their velocities), float64.

Same model and constants as ObjectTracker/byteTrack/dtypes/kalman_filter.py:40-226 (std weights 1/20 and
1/160 :52-53; initiate :55-86; predict/multi_predict :88-124,155-192; project :126-153; update via a Cholesky
solve :194-226).  The tracker keeps this 8x8 fp64 arithmetic on the host: O(#tracks) * a few hundred flops
per frame, strictly sequential per stream.
"""
import numpy as np
import scipy.linalg

W_POS = 1.0 / 20
W_VEL = 1.0 / 160

F = np.eye(8)
F[:4, 4:] = np.eye(4)          # dt = 1
H = np.eye(4, 8)


def initiate(xyah):
    mean = np.r_[xyah, np.zeros(4)]
    h = xyah[3]
    std = [2 * W_POS * h, 2 * W_POS * h, 1e-2, 2 * W_POS * h, 10 * W_VEL * h, 10 * W_VEL * h, 1e-5, 10 * W_VEL * h]
    return mean, np.diag(np.square(std))


def multi_predict(mean, cov):
    """mean [N,8], cov [N,8,8] -> predicted (operation order follows kalman_filter.py:169-192)."""
    h = mean[:, 3]
    std_pos = [W_POS * h, W_POS * h, 1e-2 * np.ones_like(h), W_POS * h]
    std_vel = [W_VEL * h, W_VEL * h, 1e-5 * np.ones_like(h), W_VEL * h]
    sqr = np.square(np.r_[std_pos, std_vel]).T
    q = np.zeros_like(cov)
    idx = np.arange(8)
    q[:, idx, idx] = sqr
    mean = np.dot(mean, F.T)
    left = np.dot(F, cov).transpose((1, 0, 2))
    return mean, np.dot(left, F.T) + q


def project(mean, cov):
    h = mean[3]
    std = [W_POS * h, W_POS * h, 1e-1, W_POS * h]
    pm = np.dot(H, mean)
    pc = np.linalg.multi_dot((H, cov, H.T))
    return pm, pc + np.diag(np.square(std))


def update(mean, cov, xyah):
    pm, pc = project(mean, cov)
    chol, lower = scipy.linalg.cho_factor(pc, lower=True, check_finite=False)
    gain = scipy.linalg.cho_solve((chol, lower), np.dot(cov, H.T).T, check_finite=False).T
    innov = xyah - pm
    return mean + np.dot(innov, gain.T), cov - np.linalg.multi_dot((gain, pc, gain.T))


def multi_update(means, covs, zs):
    """Batched `update` for N (track, measurement) pairs: same algebra as kalman_filter.py:194-226 with the 4x4 systems
    solved by one batched LAPACK call instead of N scipy Cholesky round trips (differences are at the 1e-16 level)."""
    h = means[:, 3]
    std = np.stack([W_POS * h, W_POS * h, np.full_like(h, 1e-1), W_POS * h], 1)
    pm = means[:, :4]
    pc = covs[:, :4, :4] + np.einsum("ni,ij->nij", np.square(std), np.eye(4))
    gain_t = np.linalg.solve(pc, covs[:, :4, :])            # [N,4,8] = K^T  (pc symmetric)
    innov = zs - pm
    new_means = means + np.einsum("ni,nij->nj", innov, gain_t)
    new_covs = covs - np.einsum("nia,nij,njb->nab", gain_t, pc, gain_t)
    return new_means, new_covs
