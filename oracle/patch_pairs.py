"""TEST INFRASTRUCTURE (oracle) — numpy restatement of the patch-feature-contrast triplet selection,
place_recognition/train_place_recognition.py:320-367, one (m, n) cloud pair at a time, written with the same
np.where / np.isin calls as the reference.  Only the random draws differ in mechanism: ``random.sample`` (:332) is
applied by the caller (entry order is an input) and ``np.random.choice`` (:364) is replaced by the counter-based
generator the CUDA kernel uses (splitmix64 of (seed, pair, entry, j)), so both sides pick the same negatives.
Parity unpinned by the reference (it ships no fixture for this loop)."""
import numpy as np

_M64 = (1 << 64) - 1


def _mix(z):
    z = (z + 0x9E3779B97F4A7C15) & _M64
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & _M64
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & _M64
    return z ^ (z >> 31)


def choice_index(seed, pair, entry, j, n):
    h = _mix(_mix((seed ^ (pair << 40)) & _M64) ^ ((entry << 20) | j))
    return ((h >> 32) * n) >> 32


def select_pair(m_center_indices, n_center_indices, entries, seed, pair):
    """entries: list of (idx1, near_indices2, far_list) in processing order (far_list as given to np.isin at :360).
    Returns indices1, pos_indices2, neg_indices2 (python lists, reference order)."""
    indices1, pos_indices2, neg_indices2 = [], [], []
    for k, (e_idx1, near, far) in enumerate(entries):
        idx1 = np.where(m_center_indices == e_idx1)[0].tolist()                      # :338
        if len(idx1) == 0:
            continue
        pos_idx2 = np.where(np.isin(n_center_indices, list(near)))[0].tolist()       # :344
        if len(pos_idx2) == 0:
            continue
        neg_idx2 = np.where(np.isin(n_center_indices, far))[0].tolist()              # :360
        if len(neg_idx2) == 0:
            continue
        indices1 += (np.ones(len(pos_idx2), dtype="int32") * idx1[0]).tolist()       # :363
        neg_indices2 += [neg_idx2[choice_index(seed, pair, k, j, len(neg_idx2))] for j in range(len(pos_idx2))]   # :364
        pos_indices2 += pos_idx2
    return indices1, pos_indices2, neg_indices2
